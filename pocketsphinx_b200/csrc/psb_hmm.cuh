// psb_hmm.cuh -- device-side Viterbi step for one HMM instance (register resident).
//
// One call = one hmm_vit_eval (hmm.c:787-805): s_i = score[i] + (-senscore[senid_i]), each
// target state takes the max over its (at most three) predecessors with the reference's
// exact tie order, history (and, for multiplexed HMMs, the senone-sequence id) follows the
// winner, everything is floored at WORST_SCORE, and the exit state is fed by the top two
// emitting states.  The five reference specialisations differ in which blocks are guarded
// and in how a missing skip arc is treated, so each is restated separately.
#pragma once
#include "psb_internal.cuh"

struct HmmCtxDev {
    int n_emit;
    int n_sen;
    const uint8_t *tp;        // [n_tmat][n_emit][n_emit+1]
    const uint16_t *sseq;     // [n_sseq][n_emit]
};

struct HmmReg {
    int score[PSB_HMM_MAX_NSTATE];
    int hist[PSB_HMM_MAX_NSTATE];
    int out_score, out_hist;
    int senid[PSB_HMM_MAX_NSTATE];   // senone ids (non-mpx) or per-state sseq ids (mpx)
    int best;
};

__device__ __forceinline__ int hmm_pick3(int t0, int t1, int t2, int &out)
{
    // hmm.c:257-271 and twins: returns 0 (self loop), 1 (state below), 2 (skip)
    if (t0 > t1) {
        if (t2 > t0) { out = t2; return 2; }
        out = t0; return 0;
    }
    if (t2 > t1) { out = t2; return 2; }
    out = t1; return 1;
}

#define PSB_FLOOR(s) do { if ((s) < PSB_WORST_SCORE) (s) = PSB_WORST_SCORE; } while (0)
#define PSB_RAISE(b, s) do { if ((s) > (b)) (b) = (s); } while (0)

// obs[i] = -senscore[senone of state i] (already negated), valid for non-mpx HMMs.
// hmm_vit_eval_3st_lr (hmm.c:530-607)
__device__ __forceinline__ int hmm_step_3st(HmmReg &h, const uint8_t *tp, const int (&obs)[PSB_HMM_MAX_NSTATE])
{
#define TP(i, j) (-(int)tp[(i) * 4 + (j)])
    int s0, s1, s2, s3, t0, t1, t2, best = PSB_WORST_SCORE;
    s2 = h.score[2] + obs[2];
    s1 = h.score[1] + obs[1];
    s0 = h.score[0] + obs[0];
    t2 = INT_MIN;                            // stale-t2 quirk, SURVEY A.1.5
    if (s1 > PSB_WORST_SCORE) {
        t1 = s2 + TP(2, 3);
        if (TP(1, 3) > PSB_TMAT_WORST) t2 = s1 + TP(1, 3);
        if (t1 > t2) { s3 = t1; h.out_hist = h.hist[2]; }
        else { s3 = t2; h.out_hist = h.hist[1]; }
        PSB_FLOOR(s3);
        h.out_score = s3;
        best = s3;
    }
    t0 = s2 + TP(2, 2);
    t1 = s1 + TP(1, 2);
    if (TP(0, 2) > PSB_TMAT_WORST) t2 = s0 + TP(0, 2);
    {
        int w = hmm_pick3(t0, t1, t2, s2);
        if (w == 2) h.hist[2] = h.hist[0];
        else if (w == 1) h.hist[2] = h.hist[1];
    }
    PSB_FLOOR(s2); PSB_RAISE(best, s2);
    h.score[2] = s2;
    t0 = s1 + TP(1, 1);
    t1 = s0 + TP(0, 1);
    if (t0 > t1) s1 = t0;
    else { s1 = t1; h.hist[1] = h.hist[0]; }
    PSB_FLOOR(s1); PSB_RAISE(best, s1);
    h.score[1] = s1;
    s0 = s0 + TP(0, 0);
    PSB_FLOOR(s0); PSB_RAISE(best, s0);
    h.score[0] = s0;
    h.best = best;
    return best;
#undef TP
}

// hmm_vit_eval_5st_lr (hmm.c:223-353)
__device__ __forceinline__ int hmm_step_5st(HmmReg &h, const uint8_t *tp, const int (&obs)[PSB_HMM_MAX_NSTATE])
{
#define TP(i, j) (-(int)tp[(i) * 6 + (j)])
    int s0, s1, s2, s3, s4, s5, t0, t1, t2, best = PSB_WORST_SCORE;
    s4 = h.score[4] + obs[4];
    s3 = h.score[3] + obs[3];
    if (s3 > PSB_WORST_SCORE) {
        t1 = s4 + TP(4, 5);
        t2 = s3 + TP(3, 5);
        if (t1 > t2) { s5 = t1; h.out_hist = h.hist[4]; }
        else { s5 = t2; h.out_hist = h.hist[3]; }
        PSB_FLOOR(s5);
        h.out_score = s5;
        best = s5;
    }
    s2 = h.score[2] + obs[2];
    if (s2 > PSB_WORST_SCORE) {
        int w = hmm_pick3(s4 + TP(4, 4), s3 + TP(3, 4), s2 + TP(2, 4), s4);
        if (w == 2) h.hist[4] = h.hist[2];
        else if (w == 1) h.hist[4] = h.hist[3];
        PSB_FLOOR(s4); PSB_RAISE(best, s4);
        h.score[4] = s4;
    }
    s1 = h.score[1] + obs[1];
    if (s1 > PSB_WORST_SCORE) {
        int w = hmm_pick3(s3 + TP(3, 3), s2 + TP(2, 3), s1 + TP(1, 3), s3);
        if (w == 2) h.hist[3] = h.hist[1];
        else if (w == 1) h.hist[3] = h.hist[2];
        PSB_FLOOR(s3); PSB_RAISE(best, s3);
        h.score[3] = s3;
    }
    s0 = h.score[0] + obs[0];
    {
        int w = hmm_pick3(s2 + TP(2, 2), s1 + TP(1, 2), s0 + TP(0, 2), s2);
        if (w == 2) h.hist[2] = h.hist[0];
        else if (w == 1) h.hist[2] = h.hist[1];
        PSB_FLOOR(s2); PSB_RAISE(best, s2);
        h.score[2] = s2;
    }
    t0 = s1 + TP(1, 1);
    t1 = s0 + TP(0, 1);
    if (t0 > t1) s1 = t0;
    else { s1 = t1; h.hist[1] = h.hist[0]; }
    PSB_FLOOR(s1); PSB_RAISE(best, s1);
    h.score[1] = s1;
    s0 = s0 + TP(0, 0);
    PSB_FLOOR(s0); PSB_RAISE(best, s0);
    h.score[0] = s0;
    h.best = best;
    return best;
#undef TP
}

// Multiplexed HMMs: senid[] holds one senone-sequence id per state (BAD_SSID = empty); the
// observation of state st is -senscore[sseq[senid[st]][st]] (hmm.h:199-209).
__device__ __forceinline__ int mpx_obs(const HmmCtxDev &c, const int16_t *senscr, const HmmReg &h, int st)
{
    return -(int)senscr[c.sseq[(size_t)h.senid[st] * c.n_emit + st]];
}

// hmm_vit_eval_3st_lr_mpx (hmm.c:610-706)
__device__ __forceinline__ int hmm_step_3st_mpx(HmmReg &h, const HmmCtxDev &c, const uint8_t *tp, const int16_t *senscr)
{
#define TP(i, j) (-(int)tp[(i) * 4 + (j)])
    int s0, s1, s2, s3, t0, t1, t2, best, w;
    t2 = INT_MIN;
    if (h.senid[2] == PSB_BAD_SSID) s2 = t1 = PSB_WORST_SCORE;
    else { s2 = h.score[2] + mpx_obs(c, senscr, h, 2); t1 = s2 + TP(2, 3); }
    if (h.senid[1] == PSB_BAD_SSID) s1 = t2 = PSB_WORST_SCORE;
    else {
        s1 = h.score[1] + mpx_obs(c, senscr, h, 1);
        if (TP(1, 3) > PSB_TMAT_WORST) t2 = s1 + TP(1, 3);
    }
    if (t1 > t2) { s3 = t1; h.out_hist = h.hist[2]; }
    else { s3 = t2; h.out_hist = h.hist[1]; }
    PSB_FLOOR(s3);
    h.out_score = s3;
    best = s3;
    s0 = h.score[0] + mpx_obs(c, senscr, h, 0);
    t0 = t1 = PSB_WORST_SCORE;
    if (s2 != PSB_WORST_SCORE) t0 = s2 + TP(2, 2);
    if (s1 != PSB_WORST_SCORE) t1 = s1 + TP(1, 2);
    if (TP(0, 2) > PSB_TMAT_WORST) t2 = s0 + TP(0, 2);
    w = hmm_pick3(t0, t1, t2, s2);
    if (w == 2) { h.hist[2] = h.hist[0]; h.senid[2] = h.senid[0]; }
    else if (w == 1) { h.hist[2] = h.hist[1]; h.senid[2] = h.senid[1]; }
    PSB_FLOOR(s2); PSB_RAISE(best, s2);
    h.score[2] = s2;
    t0 = PSB_WORST_SCORE;
    if (s1 != PSB_WORST_SCORE) t0 = s1 + TP(1, 1);
    t1 = s0 + TP(0, 1);
    if (t0 > t1) s1 = t0;
    else { s1 = t1; h.hist[1] = h.hist[0]; h.senid[1] = h.senid[0]; }
    PSB_FLOOR(s1); PSB_RAISE(best, s1);
    h.score[1] = s1;
    s0 += TP(0, 0);
    PSB_FLOOR(s0); PSB_RAISE(best, s0);
    h.score[0] = s0;
    h.best = best;
    return best;
#undef TP
}

// hmm_vit_eval_5st_lr_mpx (hmm.c:356-525)
__device__ __forceinline__ int hmm_step_5st_mpx(HmmReg &h, const HmmCtxDev &c, const uint8_t *tp, const int16_t *senscr)
{
#define TP(i, j) (-(int)tp[(i) * 6 + (j)])
    int s0, s1, s2, s3, s4, s5, t0, t1, t2, best, w;
    if (h.senid[4] == PSB_BAD_SSID) s4 = t1 = PSB_WORST_SCORE;
    else { s4 = h.score[4] + mpx_obs(c, senscr, h, 4); t1 = s4 + TP(4, 5); }
    if (h.senid[3] == PSB_BAD_SSID) s3 = t2 = PSB_WORST_SCORE;
    else { s3 = h.score[3] + mpx_obs(c, senscr, h, 3); t2 = s3 + TP(3, 5); }
    if (t1 > t2) { s5 = t1; h.out_hist = h.hist[4]; }
    else { s5 = t2; h.out_hist = h.hist[3]; }
    PSB_FLOOR(s5);
    h.out_score = s5;
    best = s5;

    if (h.senid[2] == PSB_BAD_SSID) s2 = t2 = PSB_WORST_SCORE;
    else { s2 = h.score[2] + mpx_obs(c, senscr, h, 2); t2 = s2 + TP(2, 4); }
    t0 = t1 = PSB_WORST_SCORE;
    if (s4 != PSB_WORST_SCORE) t0 = s4 + TP(4, 4);
    if (s3 != PSB_WORST_SCORE) t1 = s3 + TP(3, 4);
    w = hmm_pick3(t0, t1, t2, s4);
    if (w == 2) { h.hist[4] = h.hist[2]; h.senid[4] = h.senid[2]; }
    else if (w == 1) { h.hist[4] = h.hist[3]; h.senid[4] = h.senid[3]; }
    PSB_FLOOR(s4); PSB_RAISE(best, s4);
    h.score[4] = s4;

    if (h.senid[1] == PSB_BAD_SSID) s1 = t2 = PSB_WORST_SCORE;
    else { s1 = h.score[1] + mpx_obs(c, senscr, h, 1); t2 = s1 + TP(1, 3); }
    t0 = t1 = PSB_WORST_SCORE;
    if (s3 != PSB_WORST_SCORE) t0 = s3 + TP(3, 3);
    if (s2 != PSB_WORST_SCORE) t1 = s2 + TP(2, 3);
    w = hmm_pick3(t0, t1, t2, s3);
    if (w == 2) { h.hist[3] = h.hist[1]; h.senid[3] = h.senid[1]; }
    else if (w == 1) { h.hist[3] = h.hist[2]; h.senid[3] = h.senid[2]; }
    PSB_FLOOR(s3); PSB_RAISE(best, s3);
    h.score[3] = s3;

    s0 = h.score[0] + mpx_obs(c, senscr, h, 0);
    t0 = t1 = PSB_WORST_SCORE;
    if (s2 != PSB_WORST_SCORE) t0 = s2 + TP(2, 2);
    if (s1 != PSB_WORST_SCORE) t1 = s1 + TP(1, 2);
    t2 = s0 + TP(0, 2);
    w = hmm_pick3(t0, t1, t2, s2);
    if (w == 2) { h.hist[2] = h.hist[0]; h.senid[2] = h.senid[0]; }
    else if (w == 1) { h.hist[2] = h.hist[1]; h.senid[2] = h.senid[1]; }
    PSB_FLOOR(s2); PSB_RAISE(best, s2);
    h.score[2] = s2;

    t0 = PSB_WORST_SCORE;
    if (s1 != PSB_WORST_SCORE) t0 = s1 + TP(1, 1);
    t1 = s0 + TP(0, 1);
    if (t0 > t1) s1 = t0;
    else { s1 = t1; h.hist[1] = h.hist[0]; h.senid[1] = h.senid[0]; }
    PSB_FLOOR(s1); PSB_RAISE(best, s1);
    h.score[1] = s1;

    s0 += TP(0, 0);
    PSB_FLOOR(s0); PSB_RAISE(best, s0);
    h.score[0] = s0;
    h.best = best;
    return best;
#undef TP
}

// hmm_vit_eval_anytopo (hmm.c:709-784) for n_emit in 1..5, mpx or not.
__device__ __forceinline__ int hmm_step_any(HmmReg &h, const HmmCtxDev &c, const uint8_t *tp, const int16_t *senscr, bool mpx)
{
    const int n = c.n_emit;
    int st[PSB_HMM_MAX_NSTATE];
#define TP(i, j) (-(int)tp[(i) * (n + 1) + (j)])
#pragma unroll
    for (int from = 0; from < PSB_HMM_MAX_NSTATE; ++from) {
        if (from < n) {
            int sid;
            if (mpx)
                sid = h.senid[from] == PSB_BAD_SSID ? PSB_BAD_SSID : c.sseq[(size_t)h.senid[from] * n + from];
            else
                sid = h.senid[from];
            const int o = sid == PSB_BAD_SSID ? PSB_WORST_SCORE : -(int)senscr[sid];
            int v = h.score[from] + o;
            if (from > 0 && v < PSB_WORST_SCORE) v = PSB_WORST_SCORE;
            st[from] = v;
        }
    }
    int scr = PSB_WORST_SCORE, bestfrom = -1, nscr, best;
#pragma unroll
    for (int from = PSB_HMM_MAX_NSTATE - 1; from >= 0; --from)
        if (from < n && TP(from, n) > PSB_TMAT_WORST && (nscr = st[from] + TP(from, n)) > scr) {
            scr = nscr;
            bestfrom = from;
        }
    h.out_score = scr;
#pragma unroll
    for (int q = 0; q < PSB_HMM_MAX_NSTATE; ++q)
        if (q == bestfrom) h.out_hist = h.hist[q];
    best = scr;
#pragma unroll
    for (int to = PSB_HMM_MAX_NSTATE - 1; to >= 0; --to) {
        if (to < n) {
            scr = TP(to, to) > PSB_TMAT_WORST ? st[to] + TP(to, to) : PSB_WORST_SCORE;
            bestfrom = -1;
#pragma unroll
            for (int from = PSB_HMM_MAX_NSTATE - 1; from >= 0; --from)
                if (from < to && TP(from, to) > PSB_TMAT_WORST && (nscr = st[from] + TP(from, to)) > scr) {
                    scr = nscr;
                    bestfrom = from;
                }
            h.score[to] = scr;
#pragma unroll
            for (int q = 0; q < PSB_HMM_MAX_NSTATE; ++q)
                if (q == bestfrom) {
                    h.hist[to] = h.hist[q];
                    if (mpx) h.senid[to] = h.senid[q];
                }
            if (best < scr) best = scr;
        }
    }
    h.best = best;
    return best;
#undef TP
}

// Dispatcher (hmm.c:786-805).  senscr: the frame's int16 scores.
__device__ __forceinline__ int hmm_step(HmmReg &h, const HmmCtxDev &c, int tmatid, bool mpx, const int16_t *senscr)
{
    const int n = c.n_emit;
    const uint8_t *tp = c.tp + (size_t)tmatid * n * (n + 1);
    if (!mpx && (n == 3 || n == 5)) {
        int obs[PSB_HMM_MAX_NSTATE];
#pragma unroll
        for (int i = 0; i < PSB_HMM_MAX_NSTATE; ++i) obs[i] = i < n ? -(int)senscr[h.senid[i]] : 0;
        return n == 3 ? hmm_step_3st(h, tp, obs) : hmm_step_5st(h, tp, obs);
    }
    if (mpx && n == 3) return hmm_step_3st_mpx(h, c, tp, senscr);
    if (mpx && n == 5) return hmm_step_5st_mpx(h, c, tp, senscr);
    return hmm_step_any(h, c, tp, senscr, mpx);
}
