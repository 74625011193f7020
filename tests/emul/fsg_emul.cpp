// tests/emul/fsg_emul.cpp -- TEST INFRASTRUCTURE.  Compiles the device search's phase code
// (pocketsphinx_b200/csrc/psb_fsg_core.h, the source fsg_search_kernel is built from) for the host
// and runs it one "thread" at a time: every block-wide loop to completion, in ascending or
// (-DPSB_FSG_EMUL_REVERSE) descending thread order.  Identical history tables in both orders and
// against the reference's golden tables show the phase logic is right and does not depend on
// the order threads run in; only the HMM update comes from the oracle (the device's own hmm_step is
// covered by the GPU parity tests).  Never linked into libpsb200.so.
#define PSB_FSG_HOST_EMUL 1
#include <stdlib.h>
#include <string.h>
#include "../../pocketsphinx_b200/csrc/psb_fsg_host.h"
extern "C" {
#include "../../oracle/ps_oracle.h"
}

namespace {
struct OracleEval {
    pso_hmmctx_t ctx;
    const FsgGraph *G;
    const int32_t *ssid, *tmatid;
    int operator()(const FsgWork &W, int p)
    {
        pso_hmm_t h;
        const int N = G->n_emit, P = G->P;
        pso_hmm_init(&ctx, &h, 0, ssid[p], tmatid[p]);
        for (int s = 0; s < N; ++s) { h.score[s] = W.score[s * P + p]; h.history[s] = W.hist[s * P + p]; }
        h.out_score = W.out_score[p]; h.out_history = W.out_hist[p]; h.bestscore = W.best[p];
        const int b = pso_hmm_vit_eval(&ctx, &h);
        for (int s = 0; s < N; ++s) { W.score[s * P + p] = h.score[s]; W.hist[s * P + p] = h.history[s]; }
        W.out_score[p] = h.out_score; W.out_hist[p] = h.out_history; W.best[p] = h.bestscore;
        return b;
    }
};
}

extern "C" int32_t
fsg_emul_run(int32_t n_emit_state, const uint8_t *tp, const uint16_t *sseq,
             int32_t n_pnode, const int32_t *pn, int32_t n_state, const int32_t *roots,
             int32_t n_link, const int32_t *links, const int32_t *nulloff, const int32_t *nullarc,
             int32_t n_ci, int32_t silcipid, int32_t start_state,
             int32_t beam, int32_t pbeam, int32_t wbeam, int32_t maxhmmpf,
             const int16_t *senscr, int32_t n_sen, int32_t T, int32_t *hist_out, int32_t cap)
{
    FsgFlat flat;
    std::string err;
    if (fsg_flatten(n_pnode, pn, n_state, roots, n_link, links, nulloff, nullarc, n_ci, flat, err) != 0) {
        fprintf(stderr, "%s\n", err.c_str());
        return -1;
    }
    FsgGraph G;
    memset(&G, 0, sizeof(G));
    fsg_graph_bind(flat, flat.buf.data(), G);
    G.n_ci = n_ci; G.n_emit = n_emit_state; G.silcipid = silcipid; G.start_state = start_state;
    G.beam = beam; G.pbeam = pbeam; G.wbeam = wbeam; G.maxhmmpf = maxhmmpf;
    std::vector<int32_t> work(fsg_work_words(flat, n_emit_state), 0x5a5a5a5a);     // poison: nothing may rely on zeroed scratch
    FsgWork W;
    fsg_work_carve(work.data(), G, W);
    W.hist_out = hist_out; W.cap = cap;
    FsgScalars S;
    memset((void *)&S, 0x5a, sizeof(S));
    OracleEval ev;
    memset(&ev.ctx, 0, sizeof(ev.ctx));
    ev.ctx.n_emit_state = n_emit_state; ev.ctx.tp = tp; ev.ctx.sseq = sseq;
    ev.G = &G; ev.ssid = flat.ssid.data(); ev.tmatid = flat.tmatid.data();
    FSG_SYNC();                                              // (race-check builds: a fresh phase per run)
    fsg_start(G, W, &S);
    for (int f = 0; f < T && !S.overflow; ++f) {
        ev.ctx.senscore = senscr + (size_t)f * n_sen;
        fsg_step(G, W, &S, f, ev);
    }
    return S.overflow ? -2 : S.n_hist;
}

extern "C" long emul_race_count(void)
{
#ifdef PSB_FSG_RACECHECK
    return fsgrace::st().n_races;
#else
    return -1;
#endif
}
