// psb_hmmctx.cuh -- the HMM context object behind psb_hmmctx_t, shared by psb_hmm.cu (hmm_vit_eval,
// phone loop, alignment, keyword spotting, phone decoding) and psb_search.cu (grammar / n-gram search).
#pragma once
#include "psb_hmm.cuh"

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
    // grow-only workspace of psb_align_batch_* (token table, phone tables, results)
    int32_t *d_al_i32, *d_al_tok;
    uint16_t *d_al_senid;
    int64_t *d_al_tokoff;
    size_t al_i32_cap, al_tok_cap, al_senid_cap, al_tokoff_cap;
    cudaEvent_t al_ev[2];
    float last_align_ms;
    // grow-only device workspace of the search entry points (psb_search.cu)
    void *d_srch[10];
    size_t srch_cap[10];
};

static inline HmmCtxDev dev_ctx(const psb_hmmctx_t *c)
{
    HmmCtxDev d;
    d.n_emit = c->n_emit; d.n_sen = c->n_sen; d.tp = c->d_tp; d.sseq = c->d_sseq;
    return d;
}
