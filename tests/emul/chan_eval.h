// tests/emul/chan_eval.h -- TEST INFRASTRUCTURE: hmm_vit_eval through the oracle on channel `c` of a
// channel-indexed SoA work area (shared by ngs_emul.cpp and ngf_emul.cpp).
#pragma once
#include <string.h>
extern "C" {
#include "../../oracle/ps_oracle.h"
}

template <class GraphT, class WorkT>
struct OracleChanEval {
    pso_hmmctx_t ctx;
    const GraphT *G;
    int operator()(const WorkT &W, int c, bool mpx, int sid = -1)      // sid: index into the static tables (tmatid, senid)
    {
        if (sid < 0) sid = c;
        pso_hmm_t h;
        const int N = G->n_emit, M = G->M;
        memset(&h, 0, sizeof(h));
        h.mpx = mpx; h.n_emit_state = (uint8_t)N; h.tmatid = (int16_t)G->tmatid[sid];
        h.ssid = mpx ? PSO_BAD_SSID : 0;
        for (int s = 0; s < N; ++s) {
            h.score[s] = W.score[s * M + c]; h.history[s] = W.hist[s * M + c];
            h.senid[s] = (uint16_t)(mpx ? W.mss[s * M + c] : G->senid[(size_t)sid * N + s]);
        }
        h.out_score = W.out_score[c]; h.out_history = W.out_hist[c]; h.bestscore = W.best[c]; h.frame = W.frame[c];
        const int b = pso_hmm_vit_eval(&ctx, &h);
        for (int s = 0; s < N; ++s) {
            W.score[s * M + c] = h.score[s]; W.hist[s * M + c] = h.history[s];
            if (mpx) W.mss[s * M + c] = h.senid[s];
        }
        W.out_score[c] = h.out_score; W.out_hist[c] = h.out_history; W.best[c] = h.bestscore;
        return b;
    }
};
