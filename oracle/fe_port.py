"""oracle/fe_port.py -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference front end for one utterance (fresh stream, full-utterance
mode): fe_spch_to_frame / fe_fft_real / fe_spec_magnitude / fe_mel_spec (fe_sigproc.c:726-1242),
fe_remove_noise (fe_noise.c:270-364), fe_mel_cep + fe_dct2 / fe_spec2cep + fe_lifter
(fe_sigproc.c:1244-1349), frame counting (fe_interface.c:352-545), batch cmn (cmn.c:136-176) and
feat_1s_c_d_dd_cep2feat with replicated edges (feat.c:579-622).  Same float32 / float64 types and
operation order as the C code, organised the way the CUDA kernels are (bit-reversed load, one
independent work item per butterfly), so it doubles as a design check of psb_fe.cu.  Pinned against
the compiled reference by tests/test_fe_port.py (bit-exact: it uses the same libm log)."""
import math

import numpy as np

F32 = np.float32


def n_frames(d, n):
    if n <= 0:
        return 0
    return (1 + (n - d["frame_size"]) // d["frame_shift"] if n >= d["frame_size"] else 0) + 1


def _frame(d, pcm, k):
    fs, sh, N = d["frame_size"], d["frame_shift"], d["fft_size"]
    start = k * sh
    seg = pcm[start:start + fs].astype(np.float64)
    ln = len(seg)
    x = np.zeros(N, np.float64)
    if float(d["alpha"]) != 0.0:
        prev = np.empty(ln, np.float64)
        prev[1:] = seg[:-1]
        prev[0] = float(pcm[start - 1]) if start > 0 else 0.0
        x[:ln] = seg - prev * float(F32(d["alpha"]))
    else:
        x[:ln] = seg
    if d["remove_dc"]:
        mean = 0.0
        for i in range(fs):
            mean += x[i]
        mean /= fs
        x[:fs] -= mean
    half = fs // 2
    x[:half] = x[:half] * d["hamming"][:half]
    x[fs - half:fs] = x[fs - half:fs] * d["hamming"][:half][::-1]
    return x


def _fft_real(d, x):
    N, m = d["fft_size"], d["fft_order"]
    rev = np.array([int(format(i, "0%db" % m)[::-1], 2) for i in range(N)])
    y = np.empty_like(x)
    y[rev] = x
    x = y
    a, b = x[0::2].copy(), x[1::2].copy()
    x[0::2], x[1::2] = a + b, a - b
    ccc, sss = d["ccc"], d["sss"]
    for k in range(1, m):
        n4, n2, n1 = k - 1, k, k + 1
        for i in range(0, N, 1 << n1):
            xt = x[i]
            x[i] = xt + x[i + (1 << n2)]
            x[i + (1 << n2)] = xt - x[i + (1 << n2)]
            x[i + (1 << n2) + (1 << n4)] = -x[i + (1 << n2) + (1 << n4)]
            for j in range(1, 1 << n4):
                i1, i2, i3, i4 = i + j, i + (1 << n2) - j, i + (1 << n2) + j, i + (1 << n2) + (1 << n2) - j
                cc, ss = ccc[j << (m - n1)], sss[j << (m - n1)]
                x1, x2, x3, x4 = x[i1], x[i2], x[i3], x[i4]
                t1 = x3 * cc + x4 * ss
                t2 = x3 * ss - x4 * cc
                x[i4] = x2 - t2
                x[i3] = -x2 - t2
                x[i2] = x1 - t1
                x[i1] = x1 + t1
    return x


def mfspec(d, pcm):
    T = n_frames(d, len(pcm))
    N, nf = d["fft_size"], d["n_filt"]
    out = np.zeros((T, nf), np.float64)
    for k in range(T):
        x = _fft_real(d, _frame(d, pcm, k))
        spec = np.empty(N // 2 + 1, np.float64)
        spec[0] = x[0] * x[0]
        j = np.arange(1, N // 2 + 1)
        spec[1:] = x[j] * x[j] + x[N - j] * x[N - j]
        for w in range(nf):
            acc = 0.0
            s0, f0 = int(d["spec_start"][w]), int(d["filt_start"][w])
            for i in range(int(d["filt_width"][w])):
                acc += spec[s0 + i] * float(d["filt_coeffs"][f0 + i])
            out[k, w] = acc
    return out


def cepstra(d, mf):
    """mel spectra [T][n_filt] float64 -> cepstra [T][n_cep] float32 (noise removal, log, transform, lifter)."""
    T, nf = mf.shape
    nc = d["n_cep"]
    out = np.zeros((T, nc), np.float32)
    lp, clp, la, cla, lb, clb, lt, mut, mg, img = 0.7, 1 - 0.7, 0.995, 1 - 0.995, 0.5, 1 - 0.5, 0.85, 0.2, 20.0, 1.0 / 20
    power = noise = floor_ = peak = None
    cosn = d["mel_cosine"].astype(np.float64)
    for t in range(T):
        m = mf[t].copy()
        if d["remove_noise"]:
            if t == 0:
                power = m.copy(); noise = m / mg; floor_ = m / mg; peak = np.zeros(nf)
            power = lp * power + clp * m
            noise = np.where(power >= noise, la * noise + cla * power, lb * noise + clb * power)
            signal = power - noise
            signal = np.where(signal < 1.0, 1.0, signal)
            floor_ = np.where(signal >= floor_, la * floor_ + cla * signal, lb * floor_ + clb * signal)
            cur = signal.copy()
            peak = peak * lt
            signal = np.where(signal < lt * peak, peak * mut, signal)
            peak = np.where(cur > peak, cur, peak)
            signal = np.where(signal < floor_, floor_, signal)
            with np.errstate(all="ignore"):
                gain = np.where(signal < mg * power, signal / power, mg)
            gain = np.where(gain < img, img, gain)
            for i in range(nf):
                l1, l2 = max(i - 4, 0), min(i + 4, nf - 1)
                coef = 0.0
                for j in range(l1, l2 + 1):
                    coef += gain[j]
                m[i] = m[i] * (coef / (l2 - l1 + 1))
        lm = np.array([math.log(v + 1e-4) for v in m], np.float64)
        for i in range(nc):
            if d["transform"] == 0:
                if i == 0:
                    c = F32(lm[0] / 2)
                    for j in range(1, nf):
                        c = F32(float(c) + lm[j])
                    c = F32(float(c) / float(nf))
                else:
                    c = F32(0)
                    for j in range(nf):
                        c = F32(float(c) + (lm[j] * cosn[i, j]) * (1 if j == 0 else 2))
                    c = F32(float(c) / (float(nf) * 2))
            else:
                if i == 0:
                    c = F32(lm[0])
                    for j in range(1, nf):
                        c = F32(float(c) + lm[j])
                    c = F32(c * F32(d["sqrt_inv_2n"] if d["transform"] == 2 else d["sqrt_inv_n"]))
                else:
                    c = F32(0)
                    for j in range(nf):
                        c = F32(float(c) + lm[j] * cosn[i, j])
                    c = F32(c * F32(d["sqrt_inv_2n"]))
            if d["lifter_val"]:
                c = F32(c * d["lifter"][i])
            out[t, i] = c
    return out


def features(d, cep):
    """batch CMN + 1s_c_d_dd.  cep float32 [T][n_cep] -> (feats [T][3 n_cep], cep after CMN)."""
    T, nc = cep.shape
    cep = cep.copy()
    if d["cmn"] == 1 and T:
        s = np.zeros(nc, np.float32)
        cnt = 0
        for t in range(T):
            if cep[t, 0] < 0:
                continue
            s = (s + cep[t]).astype(np.float32)
            cnt += 1
        with np.errstate(all="ignore"):
            mean = (s / F32(cnt)).astype(np.float32)
        cep = (cep - mean).astype(np.float32)
    out = np.zeros((T, 3 * nc), np.float32)
    W = d["window"] - 1
    idx = lambda t: min(max(t, 0), T - 1)
    for t in range(T):
        out[t, :nc] = cep[t]
        out[t, nc:2 * nc] = cep[idx(t + W)] - cep[idx(t - W)]
        d1 = cep[idx(t + W + 1)] - cep[idx(t - W + 1)]
        d2 = cep[idx(t + W - 1)] - cep[idx(t - W - 1)]
        out[t, 2 * nc:] = d1 - d2
    return out, cep


def featurize(d, pcm):
    return features(d, cepstra(d, mfspec(d, np.ascontiguousarray(pcm, np.int16))))[0]
