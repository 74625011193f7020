// psb_search_warp.cu -- the search kernels bound one WARP per utterance (four utterances per CTA):
// the same phase code (psb_fsg_core.h, psb_ngs_core.h, psb_ngf_core.h) with warp-wide loops,
// __syncwarp() barriers and a shuffle scan.  Launchers only; the C ABI lives in psb_search.cu and
// selects these with PSB_SEARCH_WARP=1 in the environment.
#define PSB_SEARCH_WARP 1
#include "psb_search.cu"
