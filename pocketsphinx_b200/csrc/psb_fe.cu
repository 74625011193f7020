// psb_fe.cu -- batched front end on the device: int16 PCM -> MFCC -> CMN ->
// dynamic features, for whole batches of utterances, each treated as a fresh stream
// (ps_start_stream + ps_process_raw(full_utt), pocketsphinx.c:1073, acmod.c:528-560).
//
// Restates, operation for operation and in the reference's float32/float64 types and order:
//   fe_pre_emphasis_int16 / fe_hamming_window / fe_spch_to_frame     (fe_sigproc.c:726-853)
//   fe_fft_real (Sorensen real-valued FFT, float64)                  (:1062-1159)
//   fe_spec_magnitude / fe_mel_spec                                  (:1161-1242)
//   fe_remove_noise and its helpers                                  (fe_noise.c:109-364)
//   fe_mel_cep, fe_spec2cep / fe_dct2, fe_lifter                     (:1244-1349)
//   frame counting of fe_process_frames + fe_end_utt                 (fe_interface.c:352-545)
//   cmn (batch)                                                      (feat/cmn.c:136-200)
//   feat_1s_c_d_dd_cep2feat with replicated edges                    (feat/feat.c:579-622, 1243-1330)
// All tables (window, twiddles, mel filters, DCT cosines, lifter) are INPUTS: the host builds
// them with the reference's own init code / libm and passes them in psb_fe_desc_t.  The only
// operation that is not bit-reproducible is log(): device libm vs glibc can differ in the last
// bit of a float64, which survives the float32 rounding of the DCT sums only rarely (parity is
// asserted at 1e-4 relative as north_star allows for float paths; see tests/test_gpu_fe.py).
#include "psb_internal.cuh"

#include <string.h>

#include <algorithm>
#include <vector>

struct psb_fe_s {
    int device;
    int frame_size, frame_shift, fft_size, fft_order, n_filt, n_cep;
    int remove_dc, remove_noise, transform, lifter_val, window, cmn, n_coeffs;
    float alpha, sqrt_inv_n, sqrt_inv_2n;
    double *d_hamming, *d_ccc, *d_sss;
    int16_t *d_spec_start, *d_filt_start, *d_filt_width;
    float *d_filt_coeffs, *d_mel_cosine, *d_lifter;
    int *d_rev;                   // bit-reversal permutation [fft_size]
    cudaStream_t stream;
    cudaEvent_t ev[2];
    // workspace
    double *d_mfspec; size_t mfspec_cap;      // [frames][n_filt]
    float *d_mfcc; size_t mfcc_cap;           // [frames][n_cep]
    int16_t *d_pcm; size_t pcm_cap;
    float *d_feats; size_t feats_cap;
    int64_t *d_samp_off; int32_t *d_frame_off; int32_t *d_frame_utt; size_t utt_cap, fu_cap;
};

namespace {

constexpr int FE_MAX_FILT = 64;
constexpr int FE_MAX_CEP = 32;

struct FeDev {
    int frame_size, frame_shift, fft_size, fft_order, n_filt, n_cep;
    int remove_dc, remove_noise, transform, lifter_val;
    float alpha, sqrt_inv_n, sqrt_inv_2n;
    const double *hamming, *ccc, *sss;
    const int16_t *spec_start, *filt_start, *filt_width;
    const float *filt_coeffs, *mel_cosine, *lifter;
    const int *rev;
};

// One CTA per frame: samples -> pre-emphasis -> window -> real FFT -> power spectrum -> mel.
// frame_utt[f] = utterance of flat frame f; frame_off[u] = first flat frame of utterance u.
__global__ void __launch_bounds__(128)
fe_frame_kernel(FeDev p, const int16_t *__restrict__ pcm, const int64_t *__restrict__ samp_off,
                const int32_t *__restrict__ frame_off, const int32_t *__restrict__ frame_utt,
                double *__restrict__ mfspec)
{
    extern __shared__ double x[];             // [fft_size] frame, then [fft_size/2 + 1] power spectrum
    double *spec = x + p.fft_size;
    const int f = blockIdx.x;
    const int u = frame_utt[f];
    const int k = f - frame_off[u];           // frame index inside the utterance
    const int64_t s0 = samp_off[u], n = samp_off[u + 1] - s0;
    const int64_t start = (int64_t)k * p.frame_shift;
    int len = (int)min((int64_t)p.frame_size, n - start);
    const int16_t *in = pcm + s0 + start;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int half = p.frame_size / 2;

    // fe_spch_to_frame: pre-emphasis (prior = the sample just before the frame), zero padding,
    // Hamming window over frame_size; stored bit-reversed for the FFT below.
    if (!p.remove_dc) {
        for (int i = tid; i < p.fft_size; i += nt) {
            double v = 0.0;
            if (i < len) {
                if (p.alpha != 0.0f) {
                    const int16_t prev = i > 0 ? in[i - 1] : (start > 0 ? in[-1] : (int16_t)0);
                    v = (double)in[i] - (double)prev * (double)p.alpha;
                }
                else
                    v = (double)in[i];
            }
            if (i < half) v = v * p.hamming[i];
            else if (i >= p.frame_size - half && i < p.frame_size) v = v * p.hamming[p.frame_size - 1 - i];
            x[p.rev[i]] = v;
        }
    }
    else {
        // remove_dc needs the reference's sequential mean (fe_sigproc.c:805-812)
        for (int i = tid; i < p.fft_size; i += nt) {
            double v = 0.0;
            if (i < len) {
                if (p.alpha != 0.0f) {
                    const int16_t prev = i > 0 ? in[i - 1] : (start > 0 ? in[-1] : (int16_t)0);
                    v = (double)in[i] - (double)prev * (double)p.alpha;
                }
                else
                    v = (double)in[i];
            }
            x[i] = v;
        }
        __syncthreads();
        __shared__ double mean_s;
        if (tid == 0) {
            double mean = 0;
            for (int i = 0; i < p.frame_size; ++i) mean += x[i];
            mean_s = mean / p.frame_size;
        }
        __syncthreads();
        double keep[8];                                      // fft_size <= 8 * 128
        int c = 0;
        for (int i = tid; i < p.fft_size; i += nt) {
            double v = x[i];
            if (i < p.frame_size) v -= mean_s;
            if (i < half) v = v * p.hamming[i];
            else if (i >= p.frame_size - half && i < p.frame_size) v = v * p.hamming[p.frame_size - 1 - i];
            keep[c++] = v;
        }
        __syncthreads();
        c = 0;
        for (int i = tid; i < p.fft_size; i += nt) x[p.rev[i]] = keep[c++];
    }
    __syncthreads();

    // fe_fft_real.  Stage 0: 2-point butterflies.
    const int N = p.fft_size, m = p.fft_order;
    for (int i = 2 * tid; i < N; i += 2 * nt) {
        const double xt = x[i];
        x[i] = xt + x[i + 1];
        x[i + 1] = xt - x[i + 1];
    }
    __syncthreads();
    // Stages 1..m-1: N/4 independent work items each = (block, j)
    for (int kk = 1; kk < m; ++kk) {
        const int n4 = kk - 1, n2 = kk, n1 = kk + 1;
        const int per = 1 << n4;                             // items per block
        for (int w = tid; w < (N >> 2); w += nt) {
            const int b = w >> n4, j = w & (per - 1);
            const int i = b << n1;
            if (j == 0) {
                const double xt = x[i];
                x[i] = xt + x[i + (1 << n2)];
                x[i + (1 << n2)] = xt - x[i + (1 << n2)];
                x[i + (1 << n2) + (1 << n4)] = -x[i + (1 << n2) + (1 << n4)];
            }
            else {
                const int i1 = i + j, i2 = i + (1 << n2) - j, i3 = i + (1 << n2) + j, i4 = i + (1 << n2) + (1 << n2) - j;
                const double cc = p.ccc[j << (m - n1)], ss = p.sss[j << (m - n1)];
                const double x1 = x[i1], x2 = x[i2], x3 = x[i3], x4 = x[i4];
                const double t1 = x3 * cc + x4 * ss;
                const double t2 = x3 * ss - x4 * cc;
                x[i4] = x2 - t2;
                x[i3] = -x2 - t2;
                x[i2] = x1 - t1;
                x[i1] = x1 + t1;
            }
        }
        __syncthreads();
    }
    // fe_spec_magnitude
    for (int j = tid; j <= N / 2; j += nt)
        spec[j] = j == 0 ? x[0] * x[0] : x[j] * x[j] + x[N - j] * x[N - j];
    __syncthreads();
    // fe_mel_spec: one thread per filter, bins in ascending order
    if (tid < p.n_filt) {
        const int ss = p.spec_start[tid], fs = p.filt_start[tid], fw = p.filt_width[tid];
        double acc = 0;
        for (int i = 0; i < fw; ++i) acc += spec[ss + i] * (double)p.filt_coeffs[fs + i];
        mfspec[(size_t)f * p.n_filt + tid] = acc;
    }
}

// One CTA (64 threads) per utterance: noise removal (sequential over frames), log, cepstral
// transform, lifter; then batch CMN; then the dynamic features.
__global__ void __launch_bounds__(64)
fe_utt_kernel(FeDev p, const int32_t *__restrict__ frame_off, double *__restrict__ mfspec,
              float *__restrict__ mfcc, float *__restrict__ feats, int cmn, int window)
{
    __shared__ double gain[FE_MAX_FILT], lm[FE_MAX_FILT];
    __shared__ float mean_s[FE_MAX_CEP];
    const int u = blockIdx.x, tid = threadIdx.x;
    const int f0 = frame_off[u], T = frame_off[u + 1] - f0;
    if (T <= 0) return;
    const int nf = p.n_filt, nc = p.n_cep;
    // fe_noise.c constants (:64-75, :214-227)
    const double lambda_power = 0.7, comp_lambda_power = 1 - 0.7, lambda_a = 0.995, comp_lambda_a = 1 - 0.995,
                 lambda_b = 0.5, comp_lambda_b = 1 - 0.5, lambda_t = 0.85, mu_t = 0.2, max_gain = 20,
                 inv_max_gain = 1.0 / 20;
    double power = 0, noise = 0, floor_ = 0, peak = 0;
    for (int t = 0; t < T; ++t) {
        const size_t fr = (size_t)(f0 + t);
        double mval = tid < nf ? mfspec[fr * nf + tid] : 0.0;
        if (p.remove_noise) {
            if (tid < nf) {
                if (t == 0) {                                            // noise_stats->undefined (:290-305)
                    power = mval;
                    noise = mval / max_gain;
                    floor_ = mval / max_gain;
                    peak = 0.0;
                }
                power = lambda_power * power + comp_lambda_power * mval;
                // fe_lower_envelope(power -> noise)
                if (power >= noise) noise = lambda_a * noise + comp_lambda_a * power;
                else noise = lambda_b * noise + comp_lambda_b * power;
                double signal = power - noise;
                if (signal < 1.0) signal = 1.0;
                // fe_lower_envelope(signal -> floor)
                if (signal >= floor_) floor_ = lambda_a * floor_ + comp_lambda_a * signal;
                else floor_ = lambda_b * floor_ + comp_lambda_b * signal;
                // fe_temp_masking
                const double cur_in = signal;
                peak *= lambda_t;
                if (signal < lambda_t * peak) signal = peak * mu_t;
                if (cur_in > peak) peak = cur_in;
                if (signal < floor_) signal = floor_;
                double g;
                if (signal < max_gain * power) g = signal / power;
                else g = max_gain;
                if (g < inv_max_gain) g = inv_max_gain;
                gain[tid] = g;
            }
            __syncthreads();
            if (tid < nf) {
                // fe_weight_smooth, window of +-4 filters
                const int l1 = (tid - 4) > 0 ? (tid - 4) : 0;
                const int l2 = (tid + 4) < (nf - 1) ? (tid + 4) : (nf - 1);
                double coef = 0;
                for (int j = l1; j <= l2; ++j) coef += gain[j];
                mval = mval * (coef / (l2 - l1 + 1));
            }
        }
        if (tid < nf) lm[tid] = log(mval + 1e-4);                        // fe_mel_cep, LOG_FLOOR
        __syncthreads();
        if (tid < nc) {
            float c;
            if (p.transform == 0) {                                      // fe_spec2cep (legacy)
                if (tid == 0) {
                    c = (float)(lm[0] / 2);
                    for (int j = 1; j < nf; ++j) c = (float)((double)c + lm[j]);
                    c = (float)((double)c / (double)nf);
                }
                else {
                    c = 0.f;
                    for (int j = 0; j < nf; ++j) {
                        const int beta = j == 0 ? 1 : 2;
                        c = (float)((double)c + (lm[j] * (double)p.mel_cosine[tid * nf + j]) * beta);
                    }
                    c = (float)((double)c / ((double)nf * 2));
                }
            }
            else {                                                       // fe_dct2
                if (tid == 0) {
                    c = (float)lm[0];
                    for (int j = 1; j < nf; ++j) c = (float)((double)c + lm[j]);
                    c = __fmul_rn(c, p.transform == 2 ? p.sqrt_inv_2n : p.sqrt_inv_n);
                }
                else {
                    c = 0.f;
                    for (int j = 0; j < nf; ++j) c = (float)((double)c + lm[j] * (double)p.mel_cosine[tid * nf + j]);
                    c = __fmul_rn(c, p.sqrt_inv_2n);
                }
            }
            if (p.lifter_val) c = __fmul_rn(c, p.lifter[tid]);           // fe_lifter
            mfcc[fr * nc + tid] = c;
        }
        __syncthreads();
    }
    // cmn() batch (cmn.c:136-176): float32 running sums over frames with c0 >= 0
    if (cmn == 1) {
        __threadfence_block();
        __syncthreads();
        if (tid < nc) {
            float sum = 0.f;
            int cnt = 0;
            for (int t = 0; t < T; ++t) {
                const float *row = mfcc + (size_t)(f0 + t) * nc;
                if (row[0] < 0) continue;
                sum = __fadd_rn(sum, row[tid]);
                ++cnt;
            }
            mean_s[tid] = __fdiv_rn(sum, (float)cnt);
        }
        __syncthreads();
        for (int i = tid; i < T * nc; i += blockDim.x) {
            float *v = mfcc + (size_t)f0 * nc + i;
            *v = __fsub_rn(*v, mean_s[i % nc]);
        }
        __threadfence_block();
        __syncthreads();
    }
    // feat_1s_c_d_dd_cep2feat (feat.c:579-622); the frames before the first / after the last
    // are copies of it (feat_s2mfc2feat_live with beginutt / endutt, feat.c:1269-1300)
    if (feats) {
        const int W = window - 1;                                        // FEAT_DCEP_WIN = 2
        const int D = 3 * nc;
        for (int i = tid; i < T * nc; i += blockDim.x) {
            const int t = i / nc, c = i % nc;
            const float *base = mfcc + (size_t)f0 * nc + c;
#define CEP(tt) base[(size_t)min(max((tt), 0), T - 1) * nc]
            float *o = feats + (size_t)(f0 + t) * D;
            o[c] = CEP(t);
            o[nc + c] = __fsub_rn(CEP(t + W), CEP(t - W));
            const float d1 = __fsub_rn(CEP(t + W + 1), CEP(t - W + 1));
            const float d2 = __fsub_rn(CEP(t + W - 1), CEP(t - W - 1));
            o[2 * nc + c] = __fsub_rn(d1, d2);
#undef CEP
        }
    }
}

static FeDev dev_fe(const psb_fe_t *fe)
{
    FeDev p;
    p.frame_size = fe->frame_size; p.frame_shift = fe->frame_shift; p.fft_size = fe->fft_size; p.fft_order = fe->fft_order;
    p.n_filt = fe->n_filt; p.n_cep = fe->n_cep; p.remove_dc = fe->remove_dc; p.remove_noise = fe->remove_noise;
    p.transform = fe->transform; p.lifter_val = fe->lifter_val; p.alpha = fe->alpha;
    p.sqrt_inv_n = fe->sqrt_inv_n; p.sqrt_inv_2n = fe->sqrt_inv_2n;
    p.hamming = fe->d_hamming; p.ccc = fe->d_ccc; p.sss = fe->d_sss;
    p.spec_start = fe->d_spec_start; p.filt_start = fe->d_filt_start; p.filt_width = fe->d_filt_width;
    p.filt_coeffs = fe->d_filt_coeffs; p.mel_cosine = fe->d_mel_cosine; p.lifter = fe->d_lifter; p.rev = fe->d_rev;
    return p;
}

template <typename T>
static int up(T **dst, const T *src, size_t n)
{
    PSB_CUDA(cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) PSB_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return PSB_OK;
}

template <typename T>
static int grow(T **p, size_t *cap, size_t need)
{
    if (need <= *cap) return PSB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = need + need / 8 + 64;
    PSB_CUDA(cudaMalloc((void **)p, *cap * sizeof(T)));
    return PSB_OK;
}

}  // namespace

extern "C" void psb_fe_free(psb_fe_t *fe)
{
    if (!fe) return;
    cudaSetDevice(fe->device);
    if (fe->stream) cudaStreamSynchronize(fe->stream);
    cudaFree(fe->d_hamming); cudaFree(fe->d_ccc); cudaFree(fe->d_sss); cudaFree(fe->d_spec_start);
    cudaFree(fe->d_filt_start); cudaFree(fe->d_filt_width); cudaFree(fe->d_filt_coeffs); cudaFree(fe->d_mel_cosine);
    cudaFree(fe->d_lifter); cudaFree(fe->d_rev); cudaFree(fe->d_mfspec); cudaFree(fe->d_mfcc); cudaFree(fe->d_pcm);
    cudaFree(fe->d_feats); cudaFree(fe->d_samp_off); cudaFree(fe->d_frame_off); cudaFree(fe->d_frame_utt);
    if (fe->ev[0]) cudaEventDestroy(fe->ev[0]);
    if (fe->ev[1]) cudaEventDestroy(fe->ev[1]);
    if (fe->stream) cudaStreamDestroy(fe->stream);
    delete fe;
}

extern "C" int psb_fe_create(const psb_fe_desc_t *d, int device, psb_fe_t **out)
{
    PSB_REQUIRE(d && out, "psb_fe_create: bad argument");
    PSB_REQUIRE(d->frame_size > 1 && d->frame_shift > 0 && d->frame_size >= d->frame_shift, "psb_fe_create: bad frame size / shift");
    PSB_REQUIRE(d->fft_size == (1 << d->fft_order) && d->fft_size >= d->frame_size && d->fft_size >= 8 && d->fft_size <= 1024,
                "psb_fe_create: fft_size must be a power of two in [max(8, frame_size), 1024] (got %d)", d->fft_size);
    PSB_REQUIRE(d->n_filt > 0 && d->n_filt <= FE_MAX_FILT && d->n_cep > 0 && d->n_cep <= FE_MAX_CEP && d->n_cep <= d->n_filt,
                "psb_fe_create: need n_cep <= n_filt <= %d and n_cep <= %d", FE_MAX_FILT, FE_MAX_CEP);
    PSB_REQUIRE(d->transform >= 0 && d->transform <= 2, "psb_fe_create: transform must be 0 (legacy), 1 (dct) or 2 (htk)");
    PSB_REQUIRE(d->cmn == 0 || d->cmn == 1, "psb_fe_create: cmn must be 0 (none) or 1 (batch); live CMN is a host-side recurrence over utterances");
    PSB_REQUIRE(d->window == 3, "psb_fe_create: only the 1s_c_d_dd feature type (window 3) is built");
    PSB_REQUIRE(d->hamming && d->ccc && d->sss && d->spec_start && d->filt_start && d->filt_width && d->filt_coeffs &&
                d->mel_cosine && (d->lifter_val == 0 || d->lifter), "psb_fe_create: missing table");
    int n_coeffs = 0;
    for (int i = 0; i < d->n_filt; ++i) {
        PSB_REQUIRE(d->filt_start[i] == n_coeffs && d->filt_width[i] >= 0 && d->spec_start[i] >= 0 &&
                    d->spec_start[i] + d->filt_width[i] <= d->fft_size / 2 + 1, "psb_fe_create: mel filter %d out of range", i);
        n_coeffs += d->filt_width[i];
    }
    PSB_REQUIRE(n_coeffs == d->n_coeffs, "psb_fe_create: n_coeffs %d != sum of filter widths %d", d->n_coeffs, n_coeffs);
    PSB_CUDA(cudaSetDevice(device));
    psb_fe_t *fe = new psb_fe_t();
    fe->device = device;
    fe->frame_size = d->frame_size; fe->frame_shift = d->frame_shift; fe->fft_size = d->fft_size; fe->fft_order = d->fft_order;
    fe->n_filt = d->n_filt; fe->n_cep = d->n_cep; fe->remove_dc = d->remove_dc; fe->remove_noise = d->remove_noise;
    fe->transform = d->transform; fe->lifter_val = d->lifter_val; fe->window = d->window; fe->cmn = d->cmn;
    fe->n_coeffs = n_coeffs; fe->alpha = d->pre_emphasis_alpha; fe->sqrt_inv_n = d->sqrt_inv_n; fe->sqrt_inv_2n = d->sqrt_inv_2n;
    std::vector<int> rev((size_t)d->fft_size);
    for (int i = 0; i < d->fft_size; ++i) {
        int r = 0;
        for (int b = 0; b < d->fft_order; ++b) r |= ((i >> b) & 1) << (d->fft_order - 1 - b);
        rev[(size_t)i] = r;
    }
    int rc = up(&fe->d_hamming, d->hamming, (size_t)d->frame_size / 2);
    if (!rc) rc = up(&fe->d_ccc, d->ccc, (size_t)d->fft_size / 4);
    if (!rc) rc = up(&fe->d_sss, d->sss, (size_t)d->fft_size / 4);
    if (!rc) rc = up(&fe->d_spec_start, d->spec_start, (size_t)d->n_filt);
    if (!rc) rc = up(&fe->d_filt_start, d->filt_start, (size_t)d->n_filt);
    if (!rc) rc = up(&fe->d_filt_width, d->filt_width, (size_t)d->n_filt);
    if (!rc) rc = up(&fe->d_filt_coeffs, d->filt_coeffs, (size_t)n_coeffs);
    if (!rc) rc = up(&fe->d_mel_cosine, d->mel_cosine, (size_t)d->n_cep * d->n_filt);
    if (!rc) rc = up(&fe->d_lifter, d->lifter, d->lifter_val ? (size_t)d->n_cep : 0);
    if (!rc) rc = up(&fe->d_rev, rev.data(), rev.size());
    cudaError_t e = cudaSuccess;
    if (!rc) {
        e = cudaStreamCreateWithFlags(&fe->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreate(&fe->ev[0]);
        if (e == cudaSuccess) e = cudaEventCreate(&fe->ev[1]);
    }
    if (rc || e != cudaSuccess) {
        if (!rc) psb_set_error("psb_fe_create: %s", cudaGetErrorString(e));
        psb_fe_free(fe);
        return rc ? rc : PSB_ERR_CUDA;
    }
    *out = fe;
    return PSB_OK;
}

extern "C" int32_t psb_fe_n_frames(const psb_fe_t *fe, int64_t n_samples)
{
    // fe_process_frames (full frames) + fe_end_utt (one more from the leftover samples)
    if (!fe || n_samples <= 0) return 0;
    const int64_t full = n_samples >= fe->frame_size ? 1 + (n_samples - fe->frame_size) / fe->frame_shift : 0;
    return (int32_t)(full + 1);
}

static int fe_run(psb_fe_t *fe, const int16_t *d_pcm, const int64_t *samp_off, int32_t n_utt, float *d_feats,
                  float *d_mfcc_out, int32_t *frame_off, float *ms)
{
    std::vector<int32_t> foff((size_t)n_utt + 1);
    foff[0] = 0;
    for (int u = 0; u < n_utt; ++u) {
        PSB_REQUIRE(samp_off[u + 1] >= samp_off[u], "psb_fe: samp_off not monotone at %d", u);
        const int64_t t = (int64_t)foff[(size_t)u] + psb_fe_n_frames(fe, samp_off[u + 1] - samp_off[u]);
        PSB_REQUIRE(t < (1ll << 31), "psb_fe: more than 2^31 frames in one batch");
        foff[(size_t)u + 1] = (int32_t)t;
    }
    const int32_t total = foff[(size_t)n_utt];
    if (frame_off) memcpy(frame_off, foff.data(), foff.size() * sizeof(int32_t));
    if (ms) *ms = 0.f;
    if (total == 0) return PSB_OK;
    std::vector<int32_t> futt((size_t)total);
    for (int u = 0; u < n_utt; ++u)
        for (int32_t f = foff[(size_t)u]; f < foff[(size_t)u + 1]; ++f) futt[(size_t)f] = u;
    int rc = grow(&fe->d_mfspec, &fe->mfspec_cap, (size_t)total * fe->n_filt);
    if (!rc) rc = grow(&fe->d_mfcc, &fe->mfcc_cap, (size_t)total * fe->n_cep);
    if (!rc && (size_t)n_utt + 1 > fe->utt_cap) {
        cudaFree(fe->d_samp_off); cudaFree(fe->d_frame_off);
        fe->d_samp_off = nullptr; fe->d_frame_off = nullptr;
        fe->utt_cap = (size_t)n_utt + 1 + 64;
        PSB_CUDA(cudaMalloc((void **)&fe->d_samp_off, fe->utt_cap * 8));
        PSB_CUDA(cudaMalloc((void **)&fe->d_frame_off, fe->utt_cap * 4));
    }
    if (!rc) rc = grow(&fe->d_frame_utt, &fe->fu_cap, (size_t)total);
    if (rc) return rc;
    PSB_CUDA(cudaMemcpyAsync(fe->d_samp_off, samp_off, ((size_t)n_utt + 1) * 8, cudaMemcpyHostToDevice, fe->stream));
    PSB_CUDA(cudaMemcpyAsync(fe->d_frame_off, foff.data(), foff.size() * 4, cudaMemcpyHostToDevice, fe->stream));
    PSB_CUDA(cudaMemcpyAsync(fe->d_frame_utt, futt.data(), futt.size() * 4, cudaMemcpyHostToDevice, fe->stream));
    const FeDev p = dev_fe(fe);
    const size_t smem = ((size_t)fe->fft_size + fe->fft_size / 2 + 1) * sizeof(double);
    PSB_CUDA(cudaEventRecord(fe->ev[0], fe->stream));
    fe_frame_kernel<<<(unsigned)total, 128, smem, fe->stream>>>(p, d_pcm, fe->d_samp_off, fe->d_frame_off, fe->d_frame_utt,
                                                              fe->d_mfspec);
    PSB_LAUNCH_CHECK();
    fe_utt_kernel<<<(unsigned)n_utt, 64, 0, fe->stream>>>(p, fe->d_frame_off, fe->d_mfspec, fe->d_mfcc, d_feats, fe->cmn,
                                                         fe->window);
    PSB_LAUNCH_CHECK();
    PSB_CUDA(cudaEventRecord(fe->ev[1], fe->stream));
    if (d_mfcc_out)
        PSB_CUDA(cudaMemcpyAsync(d_mfcc_out, fe->d_mfcc, (size_t)total * fe->n_cep * 4, cudaMemcpyDeviceToDevice, fe->stream));
    PSB_CUDA(cudaStreamSynchronize(fe->stream));
    if (ms) PSB_CUDA(cudaEventElapsedTime(ms, fe->ev[0], fe->ev[1]));
    return PSB_OK;
}

extern "C" int psb_fe_process_device(psb_fe_t *fe, const int16_t *d_pcm, const int64_t *samp_off, int32_t n_utt,
                                     float *d_feats, float *d_mfcc, int32_t *frame_off, float *ms)
{
    PSB_REQUIRE(fe && samp_off && n_utt >= 0 && (d_pcm || samp_off[n_utt] == samp_off[0]), "psb_fe_process_device: bad argument");
    PSB_REQUIRE(samp_off[0] == 0, "psb_fe_process_device: samp_off[0] must be 0");
    PSB_CUDA(cudaSetDevice(fe->device));
    return fe_run(fe, d_pcm, samp_off, n_utt, d_feats, d_mfcc, frame_off, ms);
}

extern "C" int psb_fe_process_host(psb_fe_t *fe, const int16_t *pcm, const int64_t *samp_off, int32_t n_utt,
                                   float *feats, float *mfcc, int32_t *frame_off)
{
    PSB_REQUIRE(fe && samp_off && n_utt >= 0 && frame_off, "psb_fe_process_host: bad argument");
    PSB_REQUIRE(samp_off[0] == 0, "psb_fe_process_host: samp_off[0] must be 0");
    PSB_CUDA(cudaSetDevice(fe->device));
    const int64_t ns = samp_off[n_utt];
    PSB_REQUIRE(ns == 0 || pcm, "psb_fe_process_host: pcm is null");
    int64_t total = 0;
    for (int u = 0; u < n_utt; ++u) total += psb_fe_n_frames(fe, samp_off[u + 1] - samp_off[u]);
    int rc = grow(&fe->d_pcm, &fe->pcm_cap, (size_t)std::max<int64_t>(ns, 1));
    if (!rc) rc = grow(&fe->d_feats, &fe->feats_cap, (size_t)std::max<int64_t>(total, 1) * 3 * fe->n_cep);
    if (rc) return rc;
    if (ns) PSB_CUDA(cudaMemcpyAsync(fe->d_pcm, pcm, (size_t)ns * 2, cudaMemcpyHostToDevice, fe->stream));
    rc = fe_run(fe, fe->d_pcm, samp_off, n_utt, fe->d_feats, nullptr, frame_off, nullptr);
    if (rc) return rc;
    if (total && feats) PSB_CUDA(cudaMemcpy(feats, fe->d_feats, (size_t)total * 3 * fe->n_cep * 4, cudaMemcpyDeviceToHost));
    if (total && mfcc) PSB_CUDA(cudaMemcpy(mfcc, fe->d_mfcc, (size_t)total * fe->n_cep * 4, cudaMemcpyDeviceToHost));
    return PSB_OK;
}

extern "C" const float *psb_fe_device_feats(const psb_fe_t *fe)
{
    return fe ? fe->d_feats : nullptr;
}

extern "C" int32_t psb_fe_feat_dim(const psb_fe_t *fe)
{
    return fe ? 3 * fe->n_cep : 0;
}
