"""Sphinx-3 acoustic-model file writers (the data formats on the input side of the hot path).

Host-side plumbing for synthetic models: writes the files the reference's loaders read
(`means`, `variances`, `sendump`, `mixture_weights`, `transition_matrices`, text `mdef`,
`feat.params`; formats: SURVEY A.3, src/util/bio.c:188-296, src/ms_gauden.c:159-229,
src/ptm_mgau.c:457-661, src/tmat.c:178-236, src/mdef.c:515-700) and mirrors, in float64 with
the same libm calls, the precomputation those loaders apply, so that a model generated here is
the same model -- bit for bit -- after the unmodified reference has loaded it
(tests/test_s3io_vs_ref.py checks that against oracle/_ref when present).
"""
import math
import os
import struct

import numpy as np

BYTE_ORDER_MAGIC = 0x11223344


def _chksum(data: bytes, s=0):
    """bio.c:266-296 chksum_accum: rotate-left-20 and add, per 32-bit word."""
    w = np.frombuffer(data, "<u4").astype(np.uint64)
    for v in w.tolist():
        s = (((s << 20) | (s >> 12)) + v) & 0xFFFFFFFF
    return s


def _chksum_fast(data: bytes, s=0):
    # same recurrence, vectorised in blocks: s_{k+1} = rotl20(s_k) + w_k  (mod 2^32)
    w = np.frombuffer(data, "<u4")
    s = int(s)
    for v in w.tolist():
        s = (((s << 20) & 0xFFFFFFFF) | (s >> 12)) + v & 0xFFFFFFFF
    return s


def _write_s3(path, header_pairs, payload: bytes):
    with open(path, "wb") as f:
        f.write(b"s3\n")
        for k, v in header_pairs:
            f.write(("%s %s\n" % (k, v)).encode())
        f.write(b"endhdr\n")
        f.write(struct.pack("<I", BYTE_ORDER_MAGIC))
        f.write(payload)
        f.write(struct.pack("<I", _chksum_fast(payload)))


def write_gauden(path, arr, n_mgau, n_feat, n_density, featlen):
    """means / variances: int32 n_mgau, n_feat, n_density, veclen[n_feat], n_floats, then
    float32 [n_mgau][n_feat][n_density][veclen] (ms_gauden.c:159-229)."""
    arr = np.ascontiguousarray(arr, "<f4").ravel()
    hdr = struct.pack("<3i", n_mgau, n_feat, n_density) + struct.pack("<%di" % n_feat, *[int(x) for x in featlen])
    hdr += struct.pack("<i", arr.size)
    _write_s3(path, [("version", "1.0"), ("chksum0", "yes")], hdr + arr.tobytes())


def write_tmat(path, tp_float):
    """transition_matrices: int32 n_tmat, n_src, n_dst, n; float32 [n_tmat][n_src][n_dst]."""
    tp_float = np.ascontiguousarray(tp_float, "<f4")
    n_tmat, n_src, n_dst = tp_float.shape
    hdr = struct.pack("<4i", n_tmat, n_src, n_dst, tp_float.size)
    _write_s3(path, [("version", "1.0"), ("chksum0", "yes")], hdr + tp_float.tobytes())


def write_mixw(path, w):
    """mixture_weights: int32 n_sen, n_feat, n_comp, n; float32 [n_sen][n_feat][n_comp]."""
    w = np.ascontiguousarray(w, "<f4")
    n_sen, n_feat, n_comp = w.shape
    hdr = struct.pack("<4i", n_sen, n_feat, n_comp, w.size)
    _write_s3(path, [("version", "1.0"), ("chksum0", "yes")], hdr + w.tobytes())


def write_sendump(path, mixw, n_feat, n_density, n_sen, mixw_cb=None):
    """sendump (ptm_mgau.c:457-661): length-prefixed title/header strings, key/value strings
    terminated by a zero length, then rows/cols (unclustered) or the 16-byte cluster codebook
    (4-bit), then per feature, per codeword one row of n_sen bytes ((n_sen+1)/2 when 4-bit)."""
    def lps(s):
        b = s.encode() + b"\0"
        return struct.pack("<i", len(b)) + b
    four = mixw_cb is not None and len(mixw_cb) == 16
    out = lps("V6 Senone Probs, Smoothed, Normalized") + lps("synthetic model written by pocketsphinx_b200.s3io")
    out += lps("feature_count %d" % n_feat) + lps("mixture_count %d" % n_density) + lps("model_count %d" % n_sen)
    if four:
        out += lps("cluster_count 16") + lps("cluster_bits 4")
    out += struct.pack("<i", 0)
    if four:
        out += bytes(bytearray(np.asarray(mixw_cb, np.uint8).tolist()))
    else:
        out += struct.pack("<2i", n_density, n_sen)
    out += np.ascontiguousarray(mixw, np.uint8).tobytes()
    with open(path, "wb") as f:
        f.write(out)


def write_mdef_text(path, n_ci, n_emit, sen2ci, n_ci_sen, n_tmat=None):
    """Text model definition (mdef.c:515-700): CI phones P0..P{n-1} (the last is SIL, a filler)
    plus synthetic word-internal triphones that own the remaining senones so that every senone
    maps to its base phone (bin_mdef sen2cimap, needed by the PTM back-end)."""
    n_sen = len(sen2ci)
    n_tmat = n_tmat or n_ci
    names = ["P%03d" % i for i in range(n_ci - 1)] + ["SIL"]
    by_ci = [[] for _ in range(n_ci)]
    for s in range(n_ci_sen, n_sen):
        by_ci[int(sen2ci[s])].append(s)
    tri = []
    for b in range(n_ci):
        own = by_ci[b]
        k = 0
        ctx = 0
        while k < len(own):
            sens = [own[min(k + j, len(own) - 1)] for j in range(n_emit)]
            k += n_emit
            lft, rt = ctx // (n_ci - 1), ctx % (n_ci - 1)      # never SIL as context
            ctx += 1
            tri.append((b, lft, rt, sens))
    lines = ["# synthetic mdef written by pocketsphinx_b200.s3io", "0.3", "%d n_base" % n_ci, "%d n_tri" % len(tri),
             "%d n_state_map" % ((n_ci + len(tri)) * (n_emit + 1)), "%d n_tied_state" % n_sen,
             "%d n_tied_ci_state" % n_ci_sen, "%d n_tied_tmat" % n_tmat, "#",
             "# Columns definitions", "#base lft  rt p attrib tmat      ... state id's ..."]
    for i in range(n_ci):
        attr = "filler" if names[i] == "SIL" else "n/a"
        st = " ".join("%5d" % (i * n_emit + j) for j in range(n_emit))
        lines.append("%5s   -   - - %7s %4d %s    N" % (names[i], attr, i % n_tmat, st))
    for b, lft, rt, sens in tri:
        attr = "filler" if names[b] == "SIL" else "n/a"
        st = " ".join("%5d" % s for s in sens)
        lines.append("%5s %3s %3s i %7s %4d %s    N" % (names[b], names[lft], names[rt], attr, b % n_tmat, st))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return len(tri)


# ---------------------------------------------------------------------------------------
# mirrors of the loaders' arithmetic (float64 + libm, like the C code)

LOGBASE = 1.0001


def precompute_gaussians(var_raw, n_mgau, n_feat, n_density, featlen, varfloor=1e-4, logbase=LOGBASE):
    """gauden_dist_precompute (ms_gauden.c:264-308): floor the variances, det = sum over
    dims of (float)(int)(log(1/sqrt(2 pi v)) / log b), var = (float)(int)(log-domain 1/(2v))."""
    inv = 1.0 / math.log(logbase)
    v32 = np.ascontiguousarray(var_raw, np.float32).ravel().copy()
    v32[v32 < np.float32(varfloor)] = np.float32(varfloor)
    vals = v32.astype(np.float64).tolist()
    two_pi = 2.0 * math.pi
    dterm = np.array([float(int(math.log(1.0 / math.sqrt(v * two_pi)) * inv)) for v in vals], np.float32)
    pvar = np.array([float(int((1.0 / (v * 2.0)) * inv)) for v in vals], np.float32)
    det = np.zeros((n_mgau, n_feat, n_density), np.float32)
    pos = 0
    for m in range(n_mgau):
        for f in range(n_feat):
            fl = int(featlen[f])
            blk = dterm[pos:pos + n_density * fl].reshape(n_density, fl)
            acc = np.zeros(n_density, np.float32)
            for j in range(fl):                      # float32 accumulation in dimension order
                acc = (acc + blk[:, j]).astype(np.float32)
            det[m, f] = acc
            pos += n_density * fl
    return pvar, det


def quantize_tmat(tp_float, tpfloor=1e-4, logbase=LOGBASE):
    """tmat_init (tmat.c:215-236): normalise each row, floor the non-zero entries, renormalise,
    then tp = min(255, (-(int)(log(p)/log b)) >> 10); zero probabilities become 255."""
    inv = 1.0 / math.log(logbase)
    t = np.ascontiguousarray(tp_float, np.float32).copy()
    out = np.zeros(t.shape, np.uint8)
    zero = -(1 << 31) >> 2                                  # logmath zero for shift 0
    for i in range(t.shape[0]):
        for j in range(t.shape[1]):
            row = t[i, j]
            s = np.float64(0.0)
            for x in row.tolist():                            # vector_sum_norm: float64 sum
                s += x
            if s != 0.0:
                row = (row.astype(np.float64) * (1.0 / s)).astype(np.float32)
            row = np.where((row != 0) & (row.astype(np.float64) < tpfloor), np.float32(tpfloor), row)   # vector_nz_floor
            s = np.float64(0.0)
            for x in row.tolist():
                s += x
            if s != 0.0:
                row = (row.astype(np.float64) * (1.0 / s)).astype(np.float32)
            for k, p in enumerate(row.tolist()):
                lp = zero if p <= 0 else int(math.log(p) * inv)
                ltp = (-lp) >> 10
                out[i, j, k] = 255 if ltp > 255 else ltp
    return out


def quantize_mixw_ms(w, mixwfloor=1e-7, logbase=LOGBASE):
    """senone_mixw_read (ms_senone.c:135-268): per (senone, stream) normalise, floor, renormalise,
    then pdf = min(255, (-(int)(log(p)/log b) + 511) >> 10).  w: float32 [n_sen][n_feat][n_cw]."""
    inv = 1.0 / math.log(logbase)
    w = np.ascontiguousarray(w, np.float32)
    out = np.zeros(w.shape, np.uint8)
    zero = -(1 << 31) >> 2
    for i in range(w.shape[0]):
        for f in range(w.shape[1]):
            row = w[i, f]
            s = 0.0
            for x in row.tolist():
                s += x
            if s != 0.0:
                row = (row.astype(np.float64) * (1.0 / s)).astype(np.float32)
            row = np.where(row.astype(np.float64) < mixwfloor, np.float32(mixwfloor), row)      # vector_floor
            s = 0.0
            for x in row.tolist():
                s += x
            if s != 0.0:
                row = (row.astype(np.float64) * (1.0 / s)).astype(np.float32)
            for c, p in enumerate(row.tolist()):
                lp = zero if p <= 0 else int(math.log(p) * inv)
                q = -lp + 511
                out[i, f, c] = (q >> 10) if q < (255 << 10) else 255
    return out


def write_model_dir(path, *, kind, n_mgau, n_feat, n_density, featlen, mean, var_raw, tp_float, sen2ci, n_ci,
                    n_emit, n_ci_sen, mixw_q=None, mixw_cb=None, mixw_float=None, feat_params=None):
    """Write a complete acoustic-model directory the reference's acmod_init can load."""
    os.makedirs(path, exist_ok=True)
    write_gauden(os.path.join(path, "means"), mean, n_mgau, n_feat, n_density, featlen)
    write_gauden(os.path.join(path, "variances"), var_raw, n_mgau, n_feat, n_density, featlen)
    write_tmat(os.path.join(path, "transition_matrices"), tp_float)
    write_mdef_text(os.path.join(path, "mdef"), n_ci, n_emit, sen2ci, n_ci_sen, n_tmat=tp_float.shape[0])
    if mixw_q is not None:
        write_sendump(os.path.join(path, "sendump"), mixw_q, n_feat, n_density, len(sen2ci), mixw_cb)
    if mixw_float is not None:
        write_mixw(os.path.join(path, "mixture_weights"), mixw_float)
    with open(os.path.join(path, "feat.params"), "w") as f:
        f.write(feat_params or "")
