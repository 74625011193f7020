// psb_search_launch.h -- launchers of the search kernels in their two bindings (psb_search.cu: one CTA
// per utterance; psb_search_warp.cu: one warp per utterance).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "psb_hmm.cuh"
#include "psb_fsg_core.h"
#include "psb_ngs_core.h"
#include "psb_ngf_core.h"

#define PSB_SEARCH_LAUNCHERS(sfx)                                                                                          \
    void psb_fsg_launch_##sfx(cudaStream_t st, int n_utt, const int16_t *senscr, const int32_t *utt_off, HmmCtxDev c,      \
                              FsgGraph G, const uint16_t *senid, const int32_t *tmatid, int32_t *work, size_t work_words,  \
                              int32_t *hist, int cap, int32_t *n_hist);                                                    \
    void psb_ngs_launch_##sfx(cudaStream_t st, int n_utt, const int16_t *senscr, const int32_t *utt_off, HmmCtxDev c,      \
                              NgsGraph G, int32_t *work, size_t work_words, const int32_t *pen, int pl_window, int32_t *bp, \
                              int bp_cap, int32_t *bss, int bss_cap, int32_t *bp_idx, int32_t *result);                    \
    void psb_ngf_launch_##sfx(cudaStream_t st, int n_utt, const int16_t *senscr, const int32_t *utt_off, HmmCtxDev c,      \
                              NgfGraph G, int32_t *work, size_t work_words, const int32_t *bp_in, int in_cap,              \
                              const int32_t *n_in, int n_in_stride, int32_t *bp, int bp_cap, int32_t *bss, int bss_cap,    \
                              int32_t *bp_idx,                                                                             \
                              int32_t *result);                                                                            \
    void psb_exscan_launch_##sfx(int32_t *a, int n, int32_t *total);
PSB_SEARCH_LAUNCHERS(cta)
PSB_SEARCH_LAUNCHERS(warp)
