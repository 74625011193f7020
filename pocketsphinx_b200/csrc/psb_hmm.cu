// psb_hmm.cu -- batched hmm_vit_eval and the device-resident phone-loop Viterbi.
#include "psb_hmm.cuh"

#include <string.h>

#include <vector>

struct psb_hmmctx_s {
    int device;
    int n_emit, n_tmat, n_sseq, n_sen;
    uint8_t *d_tp;
    uint16_t *d_sseq;
    cudaStream_t stream;
    // staging for psb_hmm_vit_eval_batch
    psb_hmm_t *d_hmms, *h_hmms;
    size_t hmm_cap;
    int16_t *d_senscr, *h_senscr;
    int32_t *d_best, *h_best;
};

struct psb_phoneloop_s {
    psb_hmmctx_t *c;
    int n_phones, window, beam, pbeam, pip;
    double penalty_weight;
    int32_t *d_ssid, *d_tmatid;
    uint16_t *d_senid;            // [n_emit][n_phones]
    cudaStream_t stream;
    int32_t *d_flags;             // [1] pathological-regime counter
};

namespace {

__device__ __forceinline__ void load_hmm(const psb_hmm_t *p, HmmReg &h, int n)
{
#pragma unroll
    for (int i = 0; i < PSB_HMM_MAX_NSTATE; ++i) {
        h.score[i] = i < n ? p->score[i] : PSB_WORST_SCORE;
        h.hist[i] = i < n ? p->history[i] : -1;
        h.senid[i] = i < n ? p->senid[i] : PSB_BAD_SSID;
    }
    h.out_score = p->out_score;
    h.out_hist = p->out_history;
    h.best = p->bestscore;
}

__device__ __forceinline__ void store_hmm(psb_hmm_t *p, const HmmReg &h, int n, bool mpx)
{
#pragma unroll
    for (int i = 0; i < PSB_HMM_MAX_NSTATE; ++i)
        if (i < n) {
            p->score[i] = h.score[i];
            p->history[i] = h.hist[i];
            if (mpx) p->senid[i] = (uint16_t)h.senid[i];
        }
    p->out_score = h.out_score;
    p->out_history = h.out_hist;
    p->bestscore = h.best;
}

// One thread per hmm_t record (array-of-structs, as the search modules keep them).
__global__ void __launch_bounds__(256)
hmm_eval_aos_kernel(psb_hmm_t *hmms, int n, HmmCtxDev c, const int16_t *__restrict__ senscr, int *best_out)
{
    __shared__ int red[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int best = PSB_WORST_SCORE;
    if (i < n) {
        psb_hmm_t *p = hmms + i;
        HmmReg h;
        load_hmm(p, h, c.n_emit);
        const bool mpx = p->mpx != 0;
        best = hmm_step(h, c, p->tmatid, mpx, senscr);
        store_hmm(p, h, c.n_emit, mpx);
    }
    best = __reduce_max_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : PSB_WORST_SCORE;
        v = __reduce_max_sync(0xffffffffu, v);
        if (threadIdx.x == 0) atomicMax(best_out, v);
    }
}

// ---------------------------------------------------------------------------------------
// Phone loop: one CTA per utterance, HMM state in shared memory (struct of arrays), threads
// stride over the HMMs; time is the sequential loop.  All phones are non-multiplexed
// (phone_loop_search.c:98-103).
struct PlParams {
    int n_phones, window, beam, pbeam, pip;
    double penalty_weight;
};

template <typename T>
__device__ __forceinline__ T block_reduce_max_pair(T v, int idx, int *sidx, T *sval, int &out_idx)
{
    // max value, smallest index among ties (first winner in ascending order)
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T ov = __shfl_xor_sync(full, v, o);
        int oi = __shfl_xor_sync(full, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) { sval[warp] = v; sidx[warp] = idx; }
    __syncthreads();
    v = sval[0]; idx = sidx[0];
    for (int w = 1; w < nw; ++w) {
        T ov = sval[w]; int oi = sidx[w];
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    out_idx = idx;
    return v;
}

__global__ void __launch_bounds__(1024)
phoneloop_kernel(const int16_t *__restrict__ senscr, const int32_t *__restrict__ utt_off, HmmCtxDev c,
                 PlParams P, const uint16_t *__restrict__ senid_g, const int32_t *__restrict__ tmatid_g,
                 const int32_t *__restrict__ ssid_g,
                 int32_t *__restrict__ best_out, int32_t *__restrict__ pen_out,
                 psb_hmm_t *__restrict__ final_out, psb_hmm_t *__restrict__ trace_out, int32_t *flags)
{
    extern __shared__ int sm[];
    const int H = P.n_phones, N = c.n_emit, tid = threadIdx.x;
    int *score = sm;                       // [N][H]
    int *hist = score + N * H;             // [N][H]
    int *out_score = hist + N * H;         // [H]
    int *out_hist = out_score + H;         // [H]
    int *bestsc = out_hist + H;            // [H]
    int *frame = bestsc + H;               // [H]
    int *pen_buf = frame + H;              // [window][H]
    int *sval = pen_buf + P.window * H;    // [32]
    int *sidx = sval + 32;                 // [32]

    const int u = blockIdx.x;
    const long long f0 = utt_off[u];
    const int T = utt_off[u + 1] - utt_off[u];

    // phone_loop_search_start (:155-175): hmm_clear + hmm_enter(0, -1, 0)
    for (int i = tid; i < H; i += blockDim.x) {
        for (int s = 0; s < N; ++s) { score[s * H + i] = PSB_WORST_SCORE; hist[s * H + i] = -1; }
        out_score[i] = PSB_WORST_SCORE; out_hist[i] = -1; bestsc[i] = PSB_WORST_SCORE;
        score[i] = 0; hist[i] = -1; frame[i] = 0;
        for (int w = 0; w < P.window; ++w) pen_buf[w * H + i] = 0;
    }
    int best_score = 0, pen_ptr = 0;
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        const int16_t *row = senscr + (f0 + t) * c.n_sen;
        // renormalize_hmms (:177-191)
        const bool renorm = best_score + 2 * P.beam < PSB_WORST_SCORE;
        int bs = PSB_WORST_SCORE;
        // evaluate_hmms (:193-214)
        for (int i = tid; i < H; i += blockDim.x) {
            HmmReg h;
#pragma unroll
            for (int s = 0; s < PSB_HMM_MAX_NSTATE; ++s) {
                h.score[s] = s < N ? score[s * H + i] : PSB_WORST_SCORE;
                h.hist[s] = s < N ? hist[s * H + i] : -1;
                h.senid[s] = s < N ? senid_g[s * H + i] : PSB_BAD_SSID;
            }
            h.out_score = out_score[i]; h.out_hist = out_hist[i]; h.best = bestsc[i];
            if (renorm) {                                    // hmm_normalize (hmm.c:206-216)
#pragma unroll
                for (int s = 0; s < PSB_HMM_MAX_NSTATE; ++s)
                    if (s < N && h.score[s] > PSB_WORST_SCORE) h.score[s] -= best_score;
                if (h.out_score > PSB_WORST_SCORE) h.out_score -= best_score;
            }
            if (frame[i] >= t) {
                int b = hmm_step(h, c, tmatid_g[i], false, row);
                if (b > bs) bs = b;
            }
#pragma unroll
            for (int s = 0; s < PSB_HMM_MAX_NSTATE; ++s)
                if (s < N) { score[s * H + i] = h.score[s]; hist[s * H + i] = h.hist[s]; }
            out_score[i] = h.out_score; out_hist[i] = h.out_hist; bestsc[i] = h.best;
        }
        int dummy;
        bs = block_reduce_max_pair<int>(bs, 0, sidx, sval, dummy);
        best_score = bs;
        // store_scores (:216-239)
        if (P.window > 0) {
            for (int i = tid; i < H; i += blockDim.x)
                pen_buf[pen_ptr * H + i] = (int)((double)(bestsc[i] - best_score) * P.penalty_weight);
            pen_ptr = (pen_ptr + 1) % P.window;
            if (pen_out)
                for (int i = tid; i < H; i += blockDim.x) {
                    int pmax = PSB_WORST_SCORE;
                    for (int w = 0; w < P.window; ++w) pmax = max(pmax, pen_buf[w * H + i]);
                    pen_out[(f0 + t) * H + i] = pmax;
                }
        }
        // prune_hmms (:241-261)
        const int nf = t + 1;
        int thresh = best_score + P.beam;
        // phone_transition (:263-299): candidates among the survivors
        const int xthresh = best_score + P.pbeam;
        int cand = INT_MIN, cidx = 0x7fffffff;
        for (int i = tid; i < H; i += blockDim.x) {
            if (frame[i] < t) continue;
            if (bestsc[i] > thresh) {
                frame[i] = nf;
                const int ns = out_score[i] + P.pip;
                if (ns > xthresh && (ns > cand)) { cand = ns; cidx = i; }
            }
            else {                                           // hmm_clear_scores (hmm.c:167-178)
                for (int s = 0; s < N; ++s) score[s * H + i] = PSB_WORST_SCORE;
                out_score[i] = PSB_WORST_SCORE;
                bestsc[i] = PSB_WORST_SCORE;
            }
        }
        int widx;
        cand = block_reduce_max_pair<int>(cand, cidx, sidx, sval, widx);
        if (tid == 0) {
            if (best_out) best_out[f0 + t] = best_score;
            // sequential-order corner the max formulation does not cover
            if (cand != INT_MIN && PSB_WORST_SCORE + P.pip > xthresh) atomicAdd(flags, 1);
        }
        if (cand != INT_MIN) {
            const int whist = out_hist[widx];
            __syncthreads();
            for (int i = tid; i < H; i += blockDim.x)
                if (frame[i] < t || cand > score[i]) {       // hmm_enter (hmm.c:198-204)
                    score[i] = cand; hist[i] = whist; frame[i] = nf;
                }
        }
        __syncthreads();
        if (trace_out)
            for (int i = tid; i < H; i += blockDim.x) {
                psb_hmm_t *p = trace_out + (f0 + t) * H + i;
                for (int s = 0; s < PSB_HMM_MAX_NSTATE; ++s) {
                    p->score[s] = s < N ? score[s * H + i] : 0;
                    p->history[s] = s < N ? hist[s * H + i] : 0;
                    p->senid[s] = s < N ? senid_g[s * H + i] : 0;
                }
                p->ctx = nullptr; p->out_score = out_score[i]; p->out_history = out_hist[i];
                p->ssid = (uint16_t)ssid_g[i]; p->bestscore = bestsc[i]; p->tmatid = (int16_t)tmatid_g[i];
                p->frame = frame[i]; p->mpx = 0; p->n_emit_state = (uint8_t)N;
            }
    }
    if (final_out)
        for (int i = tid; i < H; i += blockDim.x) {
            psb_hmm_t *p = final_out + (size_t)u * H + i;
            for (int s = 0; s < PSB_HMM_MAX_NSTATE; ++s) {
                p->score[s] = s < N ? score[s * H + i] : 0;
                p->history[s] = s < N ? hist[s * H + i] : 0;
                p->senid[s] = s < N ? senid_g[s * H + i] : 0;
            }
            p->ctx = nullptr; p->out_score = out_score[i]; p->out_history = out_hist[i];
            p->ssid = (uint16_t)ssid_g[i]; p->bestscore = bestsc[i]; p->tmatid = (int16_t)tmatid_g[i];
            p->frame = frame[i]; p->mpx = 0; p->n_emit_state = (uint8_t)N;
        }
}

}  // namespace

// ---------------------------------------------------------------------------------------

extern "C" int psb_hmmctx_create(int32_t n_emit_state, const uint8_t *tp, int32_t n_tmat, const uint16_t *sseq,
                                 int32_t n_sseq, int32_t n_sen, int device, psb_hmmctx_t **out)
{
    PSB_REQUIRE(out && tp && n_emit_state >= 1 && n_emit_state <= PSB_HMM_MAX_NSTATE && n_tmat > 0 && n_sen > 0,
                "psb_hmmctx_create: bad argument");
    PSB_CUDA(cudaSetDevice(device));
    psb_hmmctx_t *c = new psb_hmmctx_t();
    memset(c, 0, sizeof(*c));
    c->device = device; c->n_emit = n_emit_state; c->n_tmat = n_tmat; c->n_sseq = n_sseq; c->n_sen = n_sen;
    size_t tpb = (size_t)n_tmat * n_emit_state * (n_emit_state + 1);
    cudaError_t e = cudaMalloc(&c->d_tp, tpb);
    if (e == cudaSuccess) e = cudaMemcpy(c->d_tp, tp, tpb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_sseq, std::max<size_t>(2, (size_t)n_sseq * n_emit_state * 2));
    if (e == cudaSuccess && n_sseq > 0 && sseq)
        e = cudaMemcpy(c->d_sseq, sseq, (size_t)n_sseq * n_emit_state * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_senscr, (size_t)n_sen * 2);
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_senscr, (size_t)n_sen * 2);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_best, 4);
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_best, 4);
    if (e != cudaSuccess) {
        psb_set_error("psb_hmmctx_create: %s", cudaGetErrorString(e));
        psb_hmmctx_free(c);
        return PSB_ERR_CUDA;
    }
    *out = c;
    return PSB_OK;
}

extern "C" void psb_hmmctx_free(psb_hmmctx_t *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_tp); cudaFree(c->d_sseq); cudaFree(c->d_hmms); cudaFree(c->d_senscr); cudaFree(c->d_best);
    if (c->h_hmms) cudaFreeHost(c->h_hmms);
    if (c->h_senscr) cudaFreeHost(c->h_senscr);
    if (c->h_best) cudaFreeHost(c->h_best);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static HmmCtxDev dev_ctx(const psb_hmmctx_t *c)
{
    HmmCtxDev d;
    d.n_emit = c->n_emit; d.n_sen = c->n_sen; d.tp = c->d_tp; d.sseq = c->d_sseq;
    return d;
}

static int validate_hmm(const psb_hmmctx_t *c, const psb_hmm_t *h, int i)
{
    PSB_REQUIRE(h->n_emit_state == c->n_emit, "hmm[%d].n_emit_state %d != context %d", i, h->n_emit_state, c->n_emit);
    PSB_REQUIRE(h->tmatid >= 0 && h->tmatid < c->n_tmat, "hmm[%d].tmatid %d out of range", i, h->tmatid);
    for (int s = 0; s < c->n_emit; ++s) {
        if (h->mpx)
            PSB_REQUIRE(h->senid[s] == PSB_BAD_SSID || h->senid[s] < c->n_sseq, "hmm[%d] ssid out of range", i);
        else
            PSB_REQUIRE(h->senid[s] < c->n_sen || (h->senid[s] == PSB_BAD_SSID && c->n_emit != 3 && c->n_emit != 5),
                        "hmm[%d].senid[%d] = %d out of range", i, s, h->senid[s]);
    }
    return PSB_OK;
}

static int ensure_hmm_cap(psb_hmmctx_t *c, size_t n)
{
    if (n <= c->hmm_cap) return PSB_OK;
    if (c->d_hmms) cudaFree(c->d_hmms);
    if (c->h_hmms) cudaFreeHost(c->h_hmms);
    c->d_hmms = nullptr; c->h_hmms = nullptr;
    c->hmm_cap = n + n / 2 + 256;
    PSB_CUDA(cudaMalloc(&c->d_hmms, c->hmm_cap * sizeof(psb_hmm_t)));
    PSB_CUDA(cudaMallocHost(&c->h_hmms, c->hmm_cap * sizeof(psb_hmm_t)));
    return PSB_OK;
}

static int eval_staged(psb_hmmctx_t *c, int32_t n, const int16_t *senscr, int32_t *best)
{
    memcpy(c->h_senscr, senscr, (size_t)c->n_sen * 2);
    *c->h_best = PSB_WORST_SCORE;
    PSB_CUDA(cudaMemcpyAsync(c->d_hmms, c->h_hmms, (size_t)n * sizeof(psb_hmm_t), cudaMemcpyHostToDevice, c->stream));
    PSB_CUDA(cudaMemcpyAsync(c->d_senscr, c->h_senscr, (size_t)c->n_sen * 2, cudaMemcpyHostToDevice, c->stream));
    PSB_CUDA(cudaMemcpyAsync(c->d_best, c->h_best, 4, cudaMemcpyHostToDevice, c->stream));
    hmm_eval_aos_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_hmms, n, dev_ctx(c), c->d_senscr, c->d_best);
    PSB_LAUNCH_CHECK();
    PSB_CUDA(cudaMemcpyAsync(c->h_hmms, c->d_hmms, (size_t)n * sizeof(psb_hmm_t), cudaMemcpyDeviceToHost, c->stream));
    PSB_CUDA(cudaMemcpyAsync(c->h_best, c->d_best, 4, cudaMemcpyDeviceToHost, c->stream));
    PSB_CUDA(cudaStreamSynchronize(c->stream));
    if (best) *best = *c->h_best;
    return PSB_OK;
}

extern "C" int psb_hmm_vit_eval_batch(psb_hmmctx_t *c, psb_hmm_t *hmms, int32_t n, const int16_t *senscr, int32_t *best)
{
    PSB_REQUIRE(c && n >= 0 && (n == 0 || (hmms && senscr)), "psb_hmm_vit_eval_batch: bad argument");
    if (best) *best = PSB_WORST_SCORE;
    if (n == 0) return PSB_OK;
    PSB_CUDA(cudaSetDevice(c->device));
    for (int i = 0; i < n; ++i) {
        int rc = validate_hmm(c, &hmms[i], i);
        if (rc) return rc;
    }
    int rc = ensure_hmm_cap(c, n);
    if (rc) return rc;
    memcpy(c->h_hmms, hmms, (size_t)n * sizeof(psb_hmm_t));
    rc = eval_staged(c, n, senscr, best);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {           // keep the caller's ctx pointers
        void *ctx = hmms[i].ctx;
        hmms[i] = c->h_hmms[i];
        hmms[i].ctx = ctx;
    }
    return PSB_OK;
}

extern "C" int psb_hmm_vit_eval_ptrs(psb_hmmctx_t *c, psb_hmm_t *const *hmms, int32_t n, const int16_t *senscr,
                                     int32_t *best)
{
    PSB_REQUIRE(c && n >= 0 && (n == 0 || (hmms && senscr)), "psb_hmm_vit_eval_ptrs: bad argument");
    if (best) *best = PSB_WORST_SCORE;
    if (n == 0) return PSB_OK;
    PSB_CUDA(cudaSetDevice(c->device));
    int rc = ensure_hmm_cap(c, n);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
        PSB_REQUIRE(hmms[i], "hmms[%d] is null", i);
        rc = validate_hmm(c, hmms[i], i);
        if (rc) return rc;
        c->h_hmms[i] = *hmms[i];
    }
    rc = eval_staged(c, n, senscr, best);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
        void *ctx = hmms[i]->ctx;
        *hmms[i] = c->h_hmms[i];
        hmms[i]->ctx = ctx;
    }
    return PSB_OK;
}

// ---------------------------------------------------------------------------------------

extern "C" int psb_phoneloop_create(psb_hmmctx_t *c, int32_t n_phones, const int32_t *ssid, const int32_t *tmatid,
                                    int32_t window, int32_t beam, int32_t pbeam, int32_t pip, double penalty_weight,
                                    psb_phoneloop_t **out)
{
    PSB_REQUIRE(c && out && n_phones > 0 && ssid && tmatid && window >= 0, "psb_phoneloop_create: bad argument");
    PSB_CUDA(cudaSetDevice(c->device));
    std::vector<uint16_t> sseq((size_t)c->n_sseq * c->n_emit);
    PSB_CUDA(cudaMemcpy(sseq.data(), c->d_sseq, sseq.size() * 2, cudaMemcpyDeviceToHost));
    std::vector<uint16_t> senid((size_t)c->n_emit * n_phones);
    for (int i = 0; i < n_phones; ++i) {
        PSB_REQUIRE(ssid[i] >= 0 && ssid[i] < c->n_sseq, "ssid[%d] out of range", i);
        PSB_REQUIRE(tmatid[i] >= 0 && tmatid[i] < c->n_tmat, "tmatid[%d] out of range", i);
        for (int s = 0; s < c->n_emit; ++s) {
            uint16_t v = sseq[(size_t)ssid[i] * c->n_emit + s];         // hmm_init, non-mpx (hmm.c:99-102)
            PSB_REQUIRE(v < c->n_sen, "senone id %d out of range", v);
            senid[(size_t)s * n_phones + i] = v;
        }
    }
    psb_phoneloop_t *p = new psb_phoneloop_t();
    memset(p, 0, sizeof(*p));
    p->c = c; p->n_phones = n_phones; p->window = window; p->beam = beam; p->pbeam = pbeam; p->pip = pip;
    p->penalty_weight = penalty_weight;
    cudaError_t e = cudaMalloc(&p->d_senid, senid.size() * 2);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_senid, senid.data(), senid.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_tmatid, (size_t)n_phones * 4);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_tmatid, tmatid, (size_t)n_phones * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_ssid, (size_t)n_phones * 4);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_ssid, ssid, (size_t)n_phones * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_flags, 4);
    if (e == cudaSuccess) e = cudaMemset(p->d_flags, 0, 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        psb_set_error("psb_phoneloop_create: %s", cudaGetErrorString(e));
        psb_phoneloop_free(p);
        return PSB_ERR_CUDA;
    }
    *out = p;
    return PSB_OK;
}

extern "C" void psb_phoneloop_free(psb_phoneloop_t *p)
{
    if (!p) return;
    cudaSetDevice(p->c->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    cudaFree(p->d_senid); cudaFree(p->d_tmatid); cudaFree(p->d_ssid); cudaFree(p->d_flags);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

int psb_phoneloop_launch(psb_phoneloop_t *p, const int16_t *d_senscr, const int32_t *d_utt_off, int32_t n_utt,
                         int32_t *d_best, int32_t *d_pen, psb_hmm_t *d_final, psb_hmm_t *d_trace, cudaStream_t st)
{
    const int H = p->n_phones, N = p->c->n_emit;
    size_t smem = ((size_t)(2 * N + 4 + p->window) * H + 64) * sizeof(int);
    PSB_REQUIRE(smem <= 227 * 1024, "phone loop with %d HMMs needs %zu bytes of shared memory (max 227 KB)", H, smem);
    PSB_CUDA(cudaFuncSetAttribute(phoneloop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = std::min(1024, roundup(H, 32));
    PlParams P;
    P.n_phones = H; P.window = p->window; P.beam = p->beam; P.pbeam = p->pbeam; P.pip = p->pip;
    P.penalty_weight = p->penalty_weight;
    phoneloop_kernel<<<n_utt, threads, smem, st>>>(d_senscr, d_utt_off, dev_ctx(p->c), P, p->d_senid, p->d_tmatid, p->d_ssid,
                                                   d_best, d_pen, d_final, d_trace, p->d_flags);
    PSB_LAUNCH_CHECK();
    return PSB_OK;
}

int psb_phoneloop_n_phones(const psb_phoneloop_t *p) { return p->n_phones; }

extern "C" int psb_phoneloop_run_device(psb_phoneloop_t *p, const int16_t *d_senscr, const int32_t *utt_off,
                                        int32_t n_utt, int32_t *d_best, int32_t *d_pen, psb_hmm_t *final_hmms,
                                        void *batch)
{
    PSB_REQUIRE(p && d_senscr && utt_off && n_utt >= 0, "psb_phoneloop_run_device: bad argument");
    if (n_utt == 0) return PSB_OK;
    PSB_CUDA(cudaSetDevice(p->c->device));
    cudaStream_t st = batch ? psb_batch_stream((psb_batch_t *)batch) : p->stream;
    int32_t *d_off = nullptr;
    psb_hmm_t *d_final = nullptr;
    PSB_CUDA(cudaMalloc(&d_off, (size_t)(n_utt + 1) * 4));
    PSB_CUDA(cudaMemcpyAsync(d_off, utt_off, (size_t)(n_utt + 1) * 4, cudaMemcpyHostToDevice, st));
    if (final_hmms) PSB_CUDA(cudaMalloc(&d_final, (size_t)n_utt * p->n_phones * sizeof(psb_hmm_t)));
    int rc = psb_phoneloop_launch(p, d_senscr, d_off, n_utt, d_best, d_pen, d_final, nullptr, st);
    if (!rc && final_hmms) {
        cudaError_t e = cudaMemcpyAsync(final_hmms, d_final, (size_t)n_utt * p->n_phones * sizeof(psb_hmm_t),
                                        cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) { psb_set_error("%s", cudaGetErrorString(e)); rc = PSB_ERR_CUDA; }
    }
    cudaStreamSynchronize(st);
    cudaFree(d_off);
    cudaFree(d_final);
    return rc;
}

extern "C" int psb_phoneloop_run_host(psb_phoneloop_t *p, const int16_t *senscr, const int32_t *utt_off, int32_t n_utt,
                                      int32_t *best, int32_t *pen, psb_hmm_t *hmm_trace)
{
    PSB_REQUIRE(p && senscr && utt_off && n_utt >= 0, "psb_phoneloop_run_host: bad argument");
    if (n_utt == 0) return PSB_OK;
    PSB_CUDA(cudaSetDevice(p->c->device));
    const size_t total = utt_off[n_utt], H = p->n_phones;
    int16_t *d_scr = nullptr; int32_t *d_off = nullptr, *d_best = nullptr, *d_pen = nullptr; psb_hmm_t *d_tr = nullptr;
    int rc = PSB_OK;
    cudaError_t e = cudaMalloc(&d_scr, std::max<size_t>(2, total * p->c->n_sen * 2));
    if (e == cudaSuccess) e = cudaMemcpy(d_scr, senscr, total * p->c->n_sen * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&d_off, (size_t)(n_utt + 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_off, utt_off, (size_t)(n_utt + 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && best) e = cudaMalloc(&d_best, std::max<size_t>(4, total * 4));
    if (e == cudaSuccess && pen) e = cudaMalloc(&d_pen, std::max<size_t>(4, total * H * 4));
    if (e == cudaSuccess && hmm_trace) e = cudaMalloc(&d_tr, std::max<size_t>(4, total * H * sizeof(psb_hmm_t)));
    if (e == cudaSuccess) {
        rc = psb_phoneloop_launch(p, d_scr, d_off, n_utt, d_best, d_pen, nullptr, d_tr, p->stream);
        if (!rc) e = cudaStreamSynchronize(p->stream);
        if (!rc && e == cudaSuccess && best) e = cudaMemcpy(best, d_best, total * 4, cudaMemcpyDeviceToHost);
        if (!rc && e == cudaSuccess && pen) e = cudaMemcpy(pen, d_pen, total * H * 4, cudaMemcpyDeviceToHost);
        if (!rc && e == cudaSuccess && hmm_trace)
            e = cudaMemcpy(hmm_trace, d_tr, total * H * sizeof(psb_hmm_t), cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) { psb_set_error("psb_phoneloop_run_host: %s", cudaGetErrorString(e)); rc = PSB_ERR_CUDA; }
    cudaFree(d_scr); cudaFree(d_off); cudaFree(d_best); cudaFree(d_pen); cudaFree(d_tr);
    return rc;
}
