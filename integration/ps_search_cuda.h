/* integration/ps_search_cuda.h -- see ps_search_cuda.c. */
#ifndef PS_SEARCH_CUDA_H
#define PS_SEARCH_CUDA_H

#include "fsg_search_internal.h"
#include "ngram_search.h"
#include "psb200.h"

typedef struct cuda_fsg_graph_s {
    psb_fsg_desc_t desc;        /* what psb_fsg_batch_device takes; points into the arrays below */
    int32 *pnodes, *roots, *links, *nulloff, *nullarc;
    fsg_link_t **link_ptr;      /* link id -> the reference's fsg_link_t (for cuda_fsg_import) */
    int n_link;
} cuda_fsg_graph_t;

int cuda_fsg_export(fsg_search_t *fs, cuda_fsg_graph_t *g);
void cuda_fsg_free(cuda_fsg_graph_t *g);
int cuda_fsg_import(fsg_search_t *fs, const cuda_fsg_graph_t *g, const int32 *rows, int32 n, int32 n_frames);

typedef struct cuda_ngram_graph_s {
    psb_ngram_desc_t desc;      /* what psb_ngram_fwdtree/fwdflat_batch_device take */
    int32 info[40];
    int32 *model;
    int64_t model_len;
    int32 *ci_tmat, *ci_ssid;
} cuda_ngram_graph_t;

int cuda_ngram_export(ngram_search_t *ngs, cuda_ngram_graph_t *g, int dense_lm);
void cuda_ngram_free(cuda_ngram_graph_t *g);
int cuda_ngram_import(ngram_search_t *ngs, const int32 *bp, int32 n, const int32 *bss, int32 n_bss,
                      const int32 *bp_idx, int32 n_frames);

/* the LM as sorted arrays (layout in ps_search_cuda.c); returns the int32 count needed / written */
long cuda_ngram_export_lm(ngram_search_t *ngs, int32 *out, long cap);

#endif
