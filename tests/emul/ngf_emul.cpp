// tests/emul/ngf_emul.cpp -- TEST INFRASTRUCTURE.  Host build of the device second pass's phase code
// (pocketsphinx_b200/csrc/psb_ngf_core.h); see fsg_emul.cpp / ngs_emul.cpp.
#define PSB_FSG_HOST_EMUL 1
#include <stdlib.h>
#include <string.h>
#include "../../pocketsphinx_b200/csrc/psb_ngf_host.h"
#include "chan_eval.h"

typedef OracleChanEval<NgfGraph, NgfWork> OracleEval;

extern "C" int32_t
ngf_emul_run(int32_t n_emit_state, const uint8_t *tp, int32_t n_tmat, const uint16_t *sseq, int32_t n_sseq, const int32_t *ci_tmat,
             const int32_t *ci_ssid, const int32_t *info, const int32_t *model, int64_t model_len, const int32_t *lm_arrays, int64_t lm_arrays_len,
             const int32_t *bp_in, int32_t n_bp_in,
             const int16_t *senscr, int32_t n_sen, int32_t T,
             int32_t *bp_out, int32_t bp_cap, int32_t *bss_out, int32_t bss_cap, int32_t *bss_n, int32_t *bp_idx_out)
{
    NgfFlat flat;
    std::string err;
    if (ngf_flatten(info, model, model_len, lm_arrays, lm_arrays_len, ci_tmat, ci_ssid, sseq, n_sseq, n_emit_state, n_tmat, n_sen, flat, err) != 0) {
        fprintf(stderr, "%s\n", err.c_str());
        return -1;
    }
    ngf_bind(flat, flat.buf.data());
    const NgfGraph &G = flat.G;
    const int n_cap = n_bp_in > 0 ? n_bp_in : 0;
    std::vector<int32_t> work(ngf_work_words(G, T, n_cap), 0x5a5a5a5a);
    NgfWork W;
    ngf_work_carve(work.data(), G, T, n_cap, W);
    W.bp = bp_out; W.bss = bss_out; W.bp_idx = bp_idx_out; W.bp_in = bp_in; W.n_bp_in = n_bp_in; W.bp_cap = bp_cap; W.bss_cap = bss_cap;
    NgfScalars S;
    memset((void *)&S, 0x5a, sizeof(S));
    OracleEval ev;
    memset(&ev.ctx, 0, sizeof(ev.ctx));
    ev.ctx.n_emit_state = n_emit_state; ev.ctx.tp = tp; ev.ctx.sseq = sseq; ev.G = &G;
    FSG_SYNC();                                              // (race-check builds: a fresh phase per run)
    ngf_start(G, W, &S);
    for (int f = 0; f < T && !S.stop && !S.error; ++f) {
        ev.ctx.senscore = senscr + (size_t)f * n_sen;
        ngf_step(G, W, &S, f, ev);
    }
    if (S.error) { fprintf(stderr, "ngf_emul: error %d at frame %d\n", (int)S.error, (int)S.n_done); return -1 - (int)S.error; }
    bp_idx_out[S.n_done] = S.bpidx;
    *bss_n = S.bss_head;
    return S.bpidx;
}

extern "C" long emul_race_count(void)
{
#ifdef PSB_FSG_RACECHECK
    return fsgrace::st().n_races;
#else
    return -1;
#endif
}
