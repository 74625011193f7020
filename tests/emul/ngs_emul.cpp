// tests/emul/ngs_emul.cpp -- TEST INFRASTRUCTURE.  Host build of the device first pass's phase code
// (pocketsphinx_b200/csrc/psb_ngs_core.h), run one "thread" at a time in ascending or
// (-DPSB_FSG_EMUL_REVERSE) descending order; the HMM update comes from the oracle.  See fsg_emul.cpp.
#define PSB_FSG_HOST_EMUL 1
#include <stdlib.h>
#include <string.h>
#include "../../pocketsphinx_b200/csrc/psb_ngs_host.h"
#include "chan_eval.h"

typedef OracleChanEval<NgsGraph, NgsWork> OracleEval;

extern "C" int32_t
ngs_emul_run(int32_t n_emit_state, const uint8_t *tp, int32_t n_tmat, const uint16_t *sseq, int32_t n_sseq, const int32_t *ci_tmat,
             const int32_t *info, const int32_t *model, int64_t model_len, const int32_t *lm_arrays, int64_t lm_arrays_len,
             const int16_t *senscr, int32_t n_sen, int32_t T,
             const int32_t *pen, int32_t pl_window, int32_t *bp_out, int32_t bp_cap, int32_t *bss_out, int32_t bss_cap, int32_t *bss_n, int32_t *bp_idx_out)
{
    NgsFlat flat;
    std::string err;
    if (ngs_flatten(info, model, model_len, lm_arrays, lm_arrays_len, ci_tmat, sseq, n_sseq, n_emit_state, n_tmat, n_sen, flat, err) != 0) {
        fprintf(stderr, "%s\n", err.c_str());
        return -1;
    }
    ngs_bind(flat, flat.buf.data());
    const NgsGraph &G = flat.G;
    std::vector<int32_t> work(ngs_work_words(G), 0x5a5a5a5a);
    NgsWork W;
    ngs_work_carve(work.data(), G, W);
    W.bp = bp_out; W.bss = bss_out; W.bp_idx = bp_idx_out; W.pen = pen; W.pl_window = pl_window; W.T = T; W.bp_cap = bp_cap; W.bss_cap = bss_cap;
    NgsScalars S;
    memset((void *)&S, 0x5a, sizeof(S));
    OracleEval ev;
    memset(&ev.ctx, 0, sizeof(ev.ctx));
    ev.ctx.n_emit_state = n_emit_state; ev.ctx.tp = tp; ev.ctx.sseq = sseq; ev.G = &G;
    FSG_SYNC();                                              // (race-check builds: a fresh phase per run)
    ngs_start(G, W, &S);
    for (int f = 0; f < T && !S.stop && !S.error; ++f) {
        ev.ctx.senscore = senscr + (size_t)f * n_sen;
        ngs_step(G, W, &S, f, ev);
    }
    if (S.error) { fprintf(stderr, "ngs_emul: error %d at frame %d: bpidx %d bss_head %d n_acl %d n_awl %d\n", (int)S.error, (int)S.n_done, (int)S.bpidx, (int)S.bss_head, (int)S.n_acl, (int)S.n_awl); return -1 - (int)S.error; }
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
