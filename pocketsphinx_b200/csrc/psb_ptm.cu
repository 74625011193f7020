// psb_ptm.cu -- batched PTM senone evaluation for sm_100a.
//
// Replaces, for whole batches of utterances with all senones computed (compallsen):
//   eval_topn / eval_cb / ptm_mgau_codebook_eval  (ptm_mgau.c:88-254)   -> ptm_topn_kernel
//   ptm_mgau_codebook_norm / ptm_mgau_senone_eval (ptm_mgau.c:266-403)  -> ptm_senone_kernel
//
// Work decomposition (DESIGN.md "Kernels"):
//  * The top-N list of a (codebook, stream) pair is a recurrence over time (frame t is seeded
//    by frame t-1's codewords), so time is the sequential axis and the parallel axes are
//    utterance x codebook x stream.  One *lane* owns one utterance; one CTA owns one
//    (codebook, stream) pair whose Gaussians sit in shared memory as warp-uniform records, so
//    every lane of a warp evaluates the same codeword against its own feature vector: the
//    model is read by LDS.128 broadcasts (one wavefront per 16 bytes for 32 lanes) and the
//    features come from a lane-major ("transposed") copy, one coalesced 128-byte load per
//    dimension per warp.
//  * The float accumulation is the reference's, rounding for rounding: x-mu, square,
//    times precomputed variance term, subtract -- four separately rounded operations per
//    dimension (__fsub_rn/__fmul_rn are never contracted into FMA), dimensions ascending.
//    The reference's early exit is result-neutral (d only decreases and the same float test
//    is repeated after the last dimension), so distances are computed in full
//    and the insertion scan is replayed per lane in codeword order.
//  * Senone evaluation is embarrassingly parallel over (utterance, frame): one CTA per
//    frame, one thread per senone, mixture weights gathered from the L2-resident table.
#include "psb_internal.cuh"

#include <algorithm>
#include <numeric>

namespace {

constexpr int TOPN = 4;               // kernels below are specialised for -topn 4 (the default)
constexpr int TOPN_WARPS = 4;         // warps per CTA in ptm_topn_kernel
constexpr int MAX_NDW = 8;            // up to 256 codewords per codebook

__device__ __forceinline__ int f2i_clamped(float d)
{
    // (int32)d, clamped first like ptm_mgau.c:129-132,219-222.  cvt.rzi saturates, which is
    // the same thing for d < INT_MIN; d > INT_MAX cannot occur (d <= det).
    return __float2int_rz(d);
}

template <int FL>
__device__ __forceinline__ float gau_dist(const float4 *__restrict__ r, const float (&x)[FL])
{
    constexpr int RECF = (1 + 2 * FL + 3) / 4 * 4;
    float rr[RECF];
#pragma unroll
    for (int q = 0; q < RECF / 4; ++q) {
        float4 v = r[q];
        rr[4 * q + 0] = v.x; rr[4 * q + 1] = v.y; rr[4 * q + 2] = v.z; rr[4 * q + 3] = v.w;
    }
    float d = rr[0];
#pragma unroll
    for (int j = 0; j < FL; ++j) {
        float diff = __fsub_rn(x[j], rr[1 + 2 * j]);
        float sq = __fmul_rn(diff, diff);
        float c = __fmul_rn(sq, rr[2 + 2 * j]);
        d = __fsub_rn(d, c);
    }
    return d;
}

// Lane/group tables built on the host per call (see psb_launch_ptm_batch):
//   lane_len[g*32+l]  frames of the utterance owned by lane l of group g (0 = padding lane)
//   lane_off[g*32+l]  flat frame offset of that utterance in feats / outputs
//   grp_base[g]       float offset of the group's block in featT; block is [maxT_g][D][32]
//   grp_maxT[g]
struct GroupTabs {
    const int32_t *lane_len, *lane_off, *grp_maxT;
    const long long *grp_base;
};

// feats [total][D] -> per group [t][D][32] (lane-major).  One warp per (group, t).
__global__ void __launch_bounds__(256)
transpose_feats_kernel(const float *__restrict__ feats, float *__restrict__ featT, GroupTabs tabs,
                       const long long *__restrict__ warp_base, int n_groups, int D)
{
    // warp_base[g] = first (group, t) work item of group g in a flat enumeration
    extern __shared__ float tile[];           // [warps][32][D+1]
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long item = (long long)blockIdx.x * warps + warp;
    // binary search the group of this item
    int lo = 0, hi = n_groups;                // warp_base has n_groups+1 entries
    if (item >= warp_base[n_groups]) return;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (warp_base[mid] <= item) lo = mid; else hi = mid;
    }
    const int g = lo;
    const int t = (int)(item - warp_base[g]);
    float *tl = tile + (size_t)warp * 32 * (D + 1);
    for (int l = 0; l < 32; ++l) {
        int len = tabs.lane_len[g * 32 + l];
        const float *row = feats + ((long long)tabs.lane_off[g * 32 + l] + t) * D;
        for (int d = lane; d < D; d += 32)
            tl[l * (D + 1) + d] = t < len ? row[d] : 0.f;
    }
    __syncwarp();
    float *dst = featT + tabs.grp_base[g] + (long long)t * D * 32;
    for (int d = 0; d < D; ++d)
        dst[d * 32 + lane] = tl[lane * (D + 1) + d];
}

// Top-N record written per (frame, codebook-stream pair), consumed by ptm_senone_kernel:
//   .x = best score >> 10           (ptm_mgau.c:277)
//   .y = codewords, byte j = cw_j
//   .z = byte j = min(255, (best >> 10) - (score_j >> 10))
//   .w = 0
// From these, ptm_mgau_codebook_norm's value min(96, norm - (score_j >> 10)) is
// min(96, (norm - .x) + .z byte j) exactly (norm >= .x, so saturating byte j at 255 is safe).

template <int FL>
__global__ void __launch_bounds__(TOPN_WARPS * 32, 6)
ptm_topn_kernel(const float *__restrict__ rec, const size_t *__restrict__ rec_off,
                const int32_t *__restrict__ klist, const float *__restrict__ featT, GroupTabs tabs,
                int4 *__restrict__ out, int n_groups, int nd, int n_feat, int D,
                const int32_t *__restrict__ featoff, int K, int ds_ratio)
{
    constexpr int RECF = (1 + 2 * FL + 3) / 4 * 4;
    constexpr int RECQ = RECF / 4;
    extern __shared__ float4 srec[];          // [nd][RECQ]
    const int k = klist[blockIdx.x];
    const int f = k % n_feat;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    {   // stage this pair's Gaussian records (L2 -> SMEM), 16 bytes per thread per step
        const float4 *src = reinterpret_cast<const float4 *>(rec + rec_off[k]);
        for (int i = threadIdx.x; i < nd * RECQ; i += blockDim.x)
            srec[i] = src[i];
    }
    __syncthreads();

    const int g = blockIdx.y * TOPN_WARPS + warp;
    if (g >= n_groups) return;
    const int len = tabs.lane_len[g * 32 + lane];
    const long long off = tabs.lane_off[g * 32 + lane];
    const int maxT = tabs.grp_maxT[g];
    const float *xT = featT + tabs.grp_base[g] + (long long)featoff[f] * 32 + lane;
    const int ndw = nd >> 5;

    int cw[TOPN], sc[TOPN];
#pragma unroll
    for (int i = 0; i < TOPN; ++i) { cw[i] = i; sc[i] = INT_MIN; }   // ptm_mgau.c:791-792

    float xn[FL];
#pragma unroll
    for (int j = 0; j < FL; ++j) xn[j] = maxT > 0 ? xT[j * 32] : 0.f;

    for (int t = 0; t < maxT; ++t) {
        float x[FL];
#pragma unroll
        for (int j = 0; j < FL; ++j) x[j] = xn[j];
        if (t + 1 < maxT) {
            const float *nx = xT + (long long)(t + 1) * D * 32;
#pragma unroll
            for (int j = 0; j < FL; ++j) xn[j] = nx[j * 32];
        }
        if (t >= len) continue;

        // ---- eval_topn (ptm_mgau.c:88-136): re-score last frame's codewords, stable sort ----
        unsigned mask[MAX_NDW];
#pragma unroll
        for (int w = 0; w < MAX_NDW; ++w) mask[w] = 0u;
        {
            int ncw[TOPN], nsc[TOPN];
#pragma unroll
            for (int i = 0; i < TOPN; ++i) {
                const int c = cw[i];
                const int s = f2i_clamped(gau_dist<FL>(srec + c * RECQ, x));
#pragma unroll
                for (int w = 0; w < MAX_NDW; ++w)
                    mask[w] |= ((c >> 5) == w) ? (1u << (c & 31)) : 0u;
                // insert (c, s) into the sorted prefix nsc[0..i-1]: entries with score < s move down
                int p = 0;
#pragma unroll
                for (int j = 0; j < i; ++j) p += (s > nsc[j]) ? 0 : 1;
#pragma unroll
                for (int j = TOPN - 2; j >= 0; --j)
                    if (j < i && j >= p) { nsc[j + 1] = nsc[j]; ncw[j + 1] = ncw[j]; }
#pragma unroll
                for (int j = 0; j < TOPN; ++j)
                    if (j == p) { nsc[j] = s; ncw[j] = c; }
            }
#pragma unroll
            for (int i = 0; i < TOPN; ++i) { cw[i] = ncw[i]; sc[i] = nsc[i]; }
        }

        // ---- eval_cb (ptm_mgau.c:152-226) unless this frame is skipped by -ds (:242) ----
        // mask[] is rotated one word per 32 codewords so that mask[0] is always the current
        // word (static register index, one copy of the loop body): logical word L sits at
        // physical position (L - w) & 7 while word w is being scanned.
        if (t % ds_ratio == 0) {
            float thresh = (float)sc[TOPN - 1];
            for (int w = 0; w < ndw; ++w) {
                const float4 *rw = srec + (size_t)w * 32 * RECQ;
#pragma unroll 2
                for (int cc = 0; cc < 32; ++cc) {
                    const float d = gau_dist<FL>(rw + cc * RECQ, x);
                    if (d >= thresh && !((mask[0] >> cc) & 1u)) {
                        const int c = w * 32 + cc;
                        const int s = f2i_clamped(d);
                        const int ev = cw[TOPN - 1];
                        // insertion_sort_cb (:140-149): entries with score <= s shift down
                        int p = 0;
#pragma unroll
                        for (int j = 0; j < TOPN - 1; ++j) p += (s >= sc[j]) ? 0 : 1;
#pragma unroll
                        for (int j = TOPN - 2; j >= 0; --j)
                            if (j >= p) { sc[j + 1] = sc[j]; cw[j + 1] = cw[j]; }
#pragma unroll
                        for (int j = 0; j < TOPN; ++j)
                            if (j == p) { sc[j] = s; cw[j] = c; }
                        // the evicted codeword is scannable again (it may lie ahead)
                        const int pw = ((ev >> 5) - w) & (MAX_NDW - 1);
                        const unsigned bit = 1u << (ev & 31);
#pragma unroll
                        for (int w2 = 0; w2 < MAX_NDW; ++w2)
                            mask[w2] &= ~((pw == w2) ? bit : 0u);
                        thresh = (float)sc[TOPN - 1];
                    }
                }
                const unsigned m0 = mask[0];
#pragma unroll
                for (int w2 = 0; w2 < MAX_NDW - 1; ++w2) mask[w2] = mask[w2 + 1];
                mask[MAX_NDW - 1] = m0;
            }
        }

        // ---- emit the record ----
        const int top = sc[0] >> PSB_SENSCR_SHIFT;
        unsigned cwb = 0, eb = 0;
#pragma unroll
        for (int j = 0; j < TOPN; ++j) {
            int e = top - (sc[j] >> PSB_SENSCR_SHIFT);
            e = e > 255 ? 255 : e;
            cwb |= (unsigned)cw[j] << (8 * j);
            eb |= (unsigned)e << (8 * j);
        }
        out[(off + t) * K + k] = make_int4(top, (int)cwb, (int)eb, 0);
    }
}

// fast_logmath_add (tied_mgau_common.h:111-127) on negated logs
__device__ __forceinline__ int logadd8(const uint8_t *tab, int x, int y)
{
    const int d = x - y;
    const int r = d > 0 ? y : x;
    return r - tab[d > 0 ? d : -d];
}

template <bool FOURBIT>
__global__ void __launch_bounds__(512)
ptm_senone_kernel(const int4 *__restrict__ topn, const uint8_t *__restrict__ mixw,
                  const uint8_t *__restrict__ mixw_cb, const uint16_t *__restrict__ sen2cb,
                  const uint8_t *__restrict__ logadd_tab, int16_t *__restrict__ senscr,
                  int n_sen, int n_feat, int nd, int K, int mixw_stride)
{
    extern __shared__ int smem_i[];
    int4 *recs = reinterpret_cast<int4 *>(smem_i);                 // [K]
    int *norm = smem_i + 4 * K;                                    // [n_feat] (+ pad to 8)
    int *red = norm + 8;                                           // [32]
    uint8_t *ns = reinterpret_cast<uint8_t *>(red + 32);           // [K*4]
    uint8_t *tab = ns + 4 * K;                                     // [256]
    uint8_t *cb16 = tab + 256;                                     // [16]
    int16_t *asc = reinterpret_cast<int16_t *>(cb16 + 16);         // [n_sen]
    const long long frame = blockIdx.x;
    const int tid = threadIdx.x;

    for (int i = tid; i < K; i += blockDim.x) recs[i] = topn[frame * K + i];
    if (tid < 256) tab[tid] = logadd_tab[tid];
    if (FOURBIT && tid < 16) cb16[tid] = mixw_cb[tid];
    if (tid < n_feat) norm[tid] = PSB_WORST_SCORE;                 // ptm_mgau.c:273
    __syncthreads();
    // ptm_mgau_codebook_norm (ptm_mgau.c:266-295), all codebooks active
    for (int i = tid; i < K; i += blockDim.x) atomicMax(&norm[i % n_feat], recs[i].x);
    __syncthreads();
    for (int i = tid; i < K; i += blockDim.x) {
        const int base = norm[i % n_feat] - recs[i].x;
        const unsigned eb = (unsigned)recs[i].z;
#pragma unroll
        for (int j = 0; j < TOPN; ++j) {
            int v = base + (int)((eb >> (8 * j)) & 0xff);
            ns[4 * i + j] = (uint8_t)(v > PSB_MAX_NEG_ASCR ? PSB_MAX_NEG_ASCR : v);
        }
    }
    __syncthreads();

    // ptm_mgau_senone_eval (ptm_mgau.c:327-403), compallsen
    int best = 0x7fffffff;
    for (int s = tid; s < n_sen; s += blockDim.x) {
        const int cb = sen2cb[s];
        int ascore = 0;
        for (int f = 0; f < n_feat; ++f) {
            const int i = cb * n_feat + f;
            const unsigned cwb = (unsigned)recs[i].y;
            const uint8_t *row = mixw + (size_t)f * nd * mixw_stride;
            int fden = 0;
#pragma unroll
            for (int j = 0; j < TOPN; ++j) {
                const int c = (cwb >> (8 * j)) & 0xff;
                int w;
                if (FOURBIT) {
                    int b = row[(size_t)c * mixw_stride + (s >> 1)];
                    b = (b & 1) ? b >> 4 : b & 0x0f;           // sic: ptm_mgau.c:376-377
                    w = cb16[b];
                }
                else
                    w = row[(size_t)c * mixw_stride + s];
                const int v = w + ns[4 * i + j];
                fden = j == 0 ? v : logadd8(tab, fden, v);
            }
            ascore += fden;
        }
        best = min(best, ascore);
        asc[s] = (int16_t)ascore;
    }
    // block-wide min
    best = __reduce_min_sync(0xffffffffu, best);
    if ((tid & 31) == 0) red[tid >> 5] = best;
    __syncthreads();
    if (tid < 32) {
        int v = tid < (int)(blockDim.x >> 5) ? red[tid] : 0x7fffffff;
        v = __reduce_min_sync(0xffffffffu, v);
        if (tid == 0) red[0] = v;
    }
    __syncthreads();
    best = red[0];
    int16_t *dst = senscr + frame * n_sen;
    for (int s = tid; s < n_sen; s += blockDim.x)
        dst[s] = (int16_t)(asc[s] - best);                       // ptm_mgau.c:398-400
}

template <int FL>
int launch_topn(psb_batch_t *b, const int32_t *d_klist, int n_k, const GroupTabs &tabs, int n_groups,
                const int32_t *d_featoff)
{
    psb_model_t *m = b->m;
    size_t smem = (size_t)m->n_density * rec_floats(FL) * sizeof(float);
    PSB_CUDA(cudaFuncSetAttribute(ptm_topn_kernel<FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(n_k, (n_groups + TOPN_WARPS - 1) / TOPN_WARPS);
    ptm_topn_kernel<FL><<<grid, TOPN_WARPS * 32, smem, b->stream>>>(
        m->d_rec, m->d_rec_off, d_klist, b->d_featT, tabs, b->d_topn, n_groups, m->n_density,
        m->n_feat, m->sumlen, d_featoff, m->K, m->ds_ratio);
    PSB_LAUNCH_CHECK();
    return PSB_OK;
}

}  // namespace

// Host side of one batched scoring pass.  d_feats: [total][D] on the device.
int psb_launch_ptm_batch(psb_batch_t *b, const float *d_feats, const int32_t *utt_off, int32_t n_utt,
                         int16_t *d_senscr)
{
    psb_model_t *m = b->m;
    PSB_REQUIRE(m->kind == PSB_KIND_PTM, "psb_launch_ptm_batch: model is not PTM");
    PSB_REQUIRE(m->topn == TOPN, "PTM batch kernels are built for -topn 4 (got %d)", m->topn);
    PSB_REQUIRE(m->n_density % 32 == 0 && m->n_density <= 32 * MAX_NDW,
                "PTM batch kernels need n_density in {32..256, multiple of 32} (got %d)", m->n_density);
    const long long total = utt_off[n_utt];
    PSB_REQUIRE(n_utt <= b->max_utts && total <= b->max_frames, "batch too large for this psb_batch_t");
    b->last_frames = total;
    if (total == 0 || n_utt == 0) return PSB_OK;
    const int D = m->sumlen, K = m->K;

    // ---- group utterances 32 per warp, longest first (ragged batches stay dense) ----
    std::vector<int> perm(n_utt);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int c) {
        return utt_off[a + 1] - utt_off[a] > utt_off[c + 1] - utt_off[c];
    });
    const int n_groups = (n_utt + 31) / 32;
    // table layout in one int32 buffer: lane_len[G*32] lane_off[G*32] grp_maxT[G] klist[K] featoff[8]
    //                                   then 8-byte aligned: grp_base[G] warp_base[G+1]
    size_t n32 = (size_t)n_groups * 64 + n_groups + K + PSB_MAX_FEAT;
    n32 = (n32 + 1) & ~(size_t)1;
    size_t need = n32 + 2 * (size_t)(2 * n_groups + 1);
    if (need > b->tab_cap) {
        if (b->d_tab) cudaFree(b->d_tab);
        if (b->h_tab) cudaFreeHost(b->h_tab);
        b->tab_cap = need * 2;
        PSB_CUDA(cudaMalloc(&b->d_tab, b->tab_cap * sizeof(int32_t)));
        PSB_CUDA(cudaMallocHost(&b->h_tab, b->tab_cap * sizeof(int32_t)));
    }
    int32_t *lane_len = b->h_tab, *lane_off = lane_len + n_groups * 32, *grp_maxT = lane_off + n_groups * 32;
    int32_t *klist = grp_maxT + n_groups, *featoff = klist + K;
    long long *grp_base = reinterpret_cast<long long *>(b->h_tab + n32), *warp_base = grp_base + n_groups;
    long long featT_floats = 0, items = 0;
    for (int g = 0; g < n_groups; ++g) {
        int mx = 0;
        for (int l = 0; l < 32; ++l) {
            int idx = g * 32 + l;
            if (idx < n_utt) {
                int u = perm[idx];
                lane_len[idx] = utt_off[u + 1] - utt_off[u];
                lane_off[idx] = utt_off[u];
                mx = std::max(mx, lane_len[idx]);
            }
            else { lane_len[idx] = 0; lane_off[idx] = 0; }
        }
        grp_maxT[g] = mx;
        grp_base[g] = featT_floats;
        warp_base[g] = items;
        featT_floats += (long long)mx * D * 32;
        items += mx;
    }
    warp_base[n_groups] = items;
    for (int f = 0; f < PSB_MAX_FEAT; ++f) featoff[f] = f < m->n_feat ? m->featoff[f] : 0;
    if ((size_t)featT_floats > b->featT_cap) {
        if (b->d_featT) cudaFree(b->d_featT);
        b->featT_cap = (size_t)featT_floats + (featT_floats >> 3);
        PSB_CUDA(cudaMalloc(&b->d_featT, b->featT_cap * sizeof(float)));
    }
    // k lists per distinct feature length
    std::vector<std::vector<int>> byfl;
    std::vector<int> fls;
    for (int f = 0; f < m->n_feat; ++f) {
        size_t i = std::find(fls.begin(), fls.end(), m->featlen[f]) - fls.begin();
        if (i == fls.size()) { fls.push_back(m->featlen[f]); byfl.emplace_back(); }
        for (int cb = 0; cb < m->n_mgau; ++cb) byfl[i].push_back(cb * m->n_feat + f);
    }
    {
        int pos = 0;
        for (auto &v : byfl) for (int k : v) klist[pos++] = k;
    }
    PSB_CUDA(cudaMemcpyAsync(b->d_tab, b->h_tab, need * sizeof(int32_t), cudaMemcpyHostToDevice, b->stream));
    GroupTabs tabs;
    tabs.lane_len = b->d_tab;
    tabs.lane_off = b->d_tab + n_groups * 32;
    tabs.grp_maxT = b->d_tab + n_groups * 64;
    const int32_t *d_klist = b->d_tab + n_groups * 64 + n_groups, *d_featoff = d_klist + K;
    tabs.grp_base = reinterpret_cast<const long long *>(b->d_tab + n32);
    const long long *d_warp_base = tabs.grp_base + n_groups;

    if (b->have_ev) PSB_CUDA(cudaEventRecord(b->ev[0], b->stream));
    {
        const int warps = 8;
        size_t smem = (size_t)warps * 32 * (D + 1) * sizeof(float);
        long long blocks = (items + warps - 1) / warps;
        transpose_feats_kernel<<<(unsigned)blocks, warps * 32, smem, b->stream>>>(
            d_feats, b->d_featT, tabs, d_warp_base, n_groups, D);
        PSB_LAUNCH_CHECK();
    }
    if (b->have_ev) PSB_CUDA(cudaEventRecord(b->ev[1], b->stream));
    {
        int pos = 0;
        for (size_t i = 0; i < fls.size(); ++i) {
            int n_k = (int)byfl[i].size(), rc;
            switch (fls[i]) {
#define CASE(FL) case FL: rc = launch_topn<FL>(b, d_klist + pos, n_k, tabs, n_groups, d_featoff); break;
                CASE(13) CASE(12) CASE(24) CASE(3) CASE(39) CASE(1) CASE(2) CASE(4) CASE(8) CASE(16) CASE(26) CASE(32)
#undef CASE
            default:
                psb_set_error("no ptm_topn_kernel instantiation for stream length %d", fls[i]);
                return PSB_ERR_ARG;
            }
            if (rc) return rc;
            pos += n_k;
        }
    }
    if (b->have_ev) PSB_CUDA(cudaEventRecord(b->ev[2], b->stream));
    {
        size_t smem = (size_t)K * 16 + 8 * 4 + 32 * 4 + (size_t)K * 4 + 256 + 16 + (size_t)m->n_sen * 2;
        if (m->mixw_4bit) {
            PSB_CUDA(cudaFuncSetAttribute(ptm_senone_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ptm_senone_kernel<true><<<(unsigned)total, 512, smem, b->stream>>>(
                b->d_topn, m->d_mixw, m->d_mixw_cb, m->d_sen2cb, m->d_logadd8, d_senscr, m->n_sen, m->n_feat,
                m->n_density, K, m->mixw_stride);
        }
        else {
            PSB_CUDA(cudaFuncSetAttribute(ptm_senone_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ptm_senone_kernel<false><<<(unsigned)total, 512, smem, b->stream>>>(
                b->d_topn, m->d_mixw, m->d_mixw_cb, m->d_sen2cb, m->d_logadd8, d_senscr, m->n_sen, m->n_feat,
                m->n_density, K, m->mixw_stride);
        }
        PSB_LAUNCH_CHECK();
    }
    if (b->have_ev) PSB_CUDA(cudaEventRecord(b->ev[3], b->stream));
    return PSB_OK;
}
