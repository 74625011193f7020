/* integration/ps_mgau_cuda.c -- reference-side binding for libpsb200.so.
 *
 * This is the file a PocketSphinx maintainer adds to src/ (it compiles against the reference's
 * own internal headers, like ptm_mgau.c does).  It provides a ps_mgau_t back-end whose vtable
 * (acmod.h:98-125) forwards to the C ABI of include/psb200.h:
 *
 *     frame_eval -> psb_scorer_frame_eval      transform -> psb_model_update_gaussians
 *     free       -> psb_scorer_free / psb_model_free
 *
 * Model files are still read by the reference's own loaders (gauden_init, read_sendump, ...):
 * cuda_mgau_wrap() takes an already initialised "ptm" back-end, uploads its arrays once and then
 * owns it.  The library is bound with dlopen so that the decoder keeps working (on the host
 * back-end) when no GPU library is installed.
 *
 * In this repository the file is compiled only into the test oracle (oracle/_ref/libpsref.so)
 * to prove the drop-in on a real decode; see INTEGRATION.md for the acmod_init_am hunk.
 */
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <pocketsphinx.h>

#include "acmod.h"
#include "ptm_mgau.h"
#include "s2_semi_mgau.h"
#include "ms_mgau.h"
#include "tied_mgau_common.h"
#include "util/ckd_alloc.h"

#include "psb200.h"

typedef struct cuda_mgau_s {
    ps_mgau_t base;            /* must be first: acmod touches base.frame_idx directly */
    ps_mgau_t *host;           /* the reference back-end that loaded the files (kept for transform) */
    psb_model_t *model;
    psb_scorer_t *scorer;
    gauden_t *g;               /* the wrapped back-end's Gaussians (all three keep a gauden_t) */
    int32 n_sen;
    void *dl;
    /* bound entry points */
    int (*model_create)(const psb_model_desc_t *, int, psb_model_t **);
    void (*model_free)(psb_model_t *);
    int (*model_update)(psb_model_t *, const float *, const float *, const float *);
    int (*scorer_create)(psb_model_t *, int32_t, psb_scorer_t **);
    void (*scorer_free)(psb_scorer_t *);
    int (*scorer_set_frame_idx)(psb_scorer_t *, int32_t);
    int (*scorer_frame_eval)(psb_scorer_t *, int16_t *, const uint8_t *, int32_t, const float *const *,
                             int32_t, int32_t);
    const char *(*last_error)(void);
    long n_calls;
} cuda_mgau_t;

static int cuda_mgau_frame_eval(ps_mgau_t *mg, int16 *senscr, uint8 *senone_active, int32 n_senone_active,
                                mfcc_t **feat, int32 frame, int32 compallsen);
static int cuda_mgau_transform(ps_mgau_t *mg, ps_mllr_t *mllr);
static void cuda_mgau_free(ps_mgau_t *mg);

static ps_mgaufuncs_t cuda_mgau_funcs = {
    "cuda",
    cuda_mgau_frame_eval,
    cuda_mgau_transform,
    cuda_mgau_free
};

static int
flatten_gauden(gauden_t *g, float **mean, float **var)
{
    int sumlen = 0, i, f;
    size_t n;
    float *m, *v;
    for (f = 0; f < g->n_feat; ++f) sumlen += g->featlen[f];
    n = (size_t)g->n_mgau * g->n_density * sumlen;
    *mean = m = ckd_calloc(n, sizeof(float));
    *var = v = ckd_calloc(n, sizeof(float));
    for (i = 0; i < g->n_mgau; ++i)
        for (f = 0; f < g->n_feat; ++f) {
            size_t nb = (size_t)g->n_density * g->featlen[f];
            memcpy(m, g->mean[i][f][0], nb * sizeof(float));
            memcpy(v, g->var[i][f][0], nb * sizeof(float));
            m += nb;
            v += nb;
        }
    return sumlen;
}

/* Wrap an initialised host back-end ("ptm", "s2_semi" or "ms").  Returns NULL (after E_ERROR)
 * if the GPU path is not available; the caller then simply keeps using `host`. */
ps_mgau_t *
cuda_mgau_wrap(acmod_t *acmod, ps_mgau_t *host, const char *libpath, int device)
{
    const char *name = host->vt->name;
    cuda_mgau_t *c;
    psb_model_desc_t d;
    gauden_t *g;
    float *mean = NULL, *var = NULL;
    int32 *s2c = NULL;
    uint32 *wide = NULL;
    uint8 *mixw = NULL;
    int32 n_sen, n_hist = 2;
    int f, cw, i, rc;

    memset(&d, 0, sizeof(d));
#ifdef FIXED_POINT
    /* mfcc_t is int32 (Q12) in this build: the arrays travel as 4-byte words and the library
     * switches to the FIXMUL / GMMSUB arithmetic (fe/fixpoint.h:98-100, tied_mgau_common.h:62-70). */
    if (strcmp(name, "ms") == 0) {
        E_ERROR("cuda_mgau: the ms back-end of a FIXED_POINT build is not supported\n");
        return NULL;
    }
    d.fixed_point = 1;
#endif
    if (strcmp(name, "ptm") == 0) {
        ptm_mgau_t *p = (ptm_mgau_t *)host;
        size_t row;
        g = p->g; n_sen = p->n_sen; n_hist = p->n_fast_hist;
        d.kind = PSB_KIND_PTM; d.topn = p->max_topn; d.ds_ratio = p->ds_ratio;
        row = p->mixw_cb ? (size_t)(n_sen + 1) / 2 : (size_t)n_sen;
        mixw = ckd_calloc((size_t)g->n_feat * g->n_density, row);
        for (f = 0; f < g->n_feat; ++f)
            for (cw = 0; cw < g->n_density; ++cw)
                memcpy(mixw + ((size_t)f * g->n_density + cw) * row, p->mixw[f][cw], row);
        d.mixw_cb = p->mixw_cb;
        s2c = ckd_calloc(n_sen, sizeof(*s2c));
        for (i = 0; i < n_sen; ++i) s2c[i] = p->sen2cb[i];
        d.logadd8 = (const uint8_t *)LOGMATH_TABLE(p->lmath_8b)->table;
    }
    else if (strcmp(name, "s2_semi") == 0) {
        s2_semi_mgau_t *p = (s2_semi_mgau_t *)host;
        size_t row;
        g = p->g; n_sen = p->n_sen; n_hist = p->n_topn_hist;
        d.kind = PSB_KIND_SEMI; d.topn = p->max_topn; d.ds_ratio = p->ds_ratio;
        row = p->mixw_cb ? (size_t)(n_sen + 1) / 2 : (size_t)n_sen;
        mixw = ckd_calloc((size_t)g->n_feat * g->n_density, row);
        for (f = 0; f < g->n_feat; ++f)
            for (cw = 0; cw < g->n_density; ++cw)
                memcpy(mixw + ((size_t)f * g->n_density + cw) * row, p->mixw[f][cw], row);
        d.mixw_cb = p->mixw_cb;
        s2c = ckd_calloc(n_sen, sizeof(*s2c));
        d.logadd8 = (const uint8_t *)LOGMATH_TABLE(p->lmath_8b)->table;
        d.topn_beam = p->topn_beam;
    }
    else if (strcmp(name, "ms") == 0) {
        ms_mgau_model_t *p = (ms_mgau_model_t *)host;
        senone_t *sen = p->s;
        logadd_t *t = LOGMATH_TABLE(sen->lmath);
        g = p->g; n_sen = sen->n_sen;
        d.kind = PSB_KIND_MS; d.topn = p->topn; d.aw = sen->aw;
        mixw = ckd_calloc((size_t)sen->n_sen * sen->n_feat, sen->n_cw);
        memcpy(mixw, sen->pdf[0][0], (size_t)sen->n_sen * sen->n_feat * sen->n_cw);
        s2c = ckd_calloc(n_sen, sizeof(*s2c));
        for (i = 0; i < n_sen; ++i) s2c[i] = sen->mgau[i];
        wide = ckd_calloc(t->table_size, sizeof(*wide));
        for (i = 0; i < (int)t->table_size; ++i)
            wide[i] = t->width == 1 ? ((uint8 *)t->table)[i] : t->width == 2 ? ((uint16 *)t->table)[i] : ((uint32 *)t->table)[i];
        d.logadd_ms = wide;
        d.logadd_ms_size = t->table_size;
        d.logadd_ms_zero = logmath_get_zero(sen->lmath);
    }
    else {
        E_ERROR("cuda_mgau: unknown back-end %s\n", name);
        return NULL;
    }
    c = ckd_calloc(1, sizeof(*c));
    c->dl = dlopen(libpath ? libpath : "libpsb200.so", RTLD_NOW | RTLD_LOCAL);
    if (c->dl == NULL) {
        E_ERROR("cuda_mgau: %s\n", dlerror());
        goto fail;
    }
#define BIND(field, sym) do { *(void **)&c->field = dlsym(c->dl, sym); \
        if (!c->field) { E_ERROR("cuda_mgau: missing symbol %s\n", sym); goto fail; } } while (0)
    BIND(model_create, "psb_model_create");
    BIND(model_free, "psb_model_free");
    BIND(model_update, "psb_model_update_gaussians");
    BIND(scorer_create, "psb_scorer_create");
    BIND(scorer_free, "psb_scorer_free");
    BIND(scorer_set_frame_idx, "psb_scorer_set_frame_idx");
    BIND(scorer_frame_eval, "psb_scorer_frame_eval");
    BIND(last_error, "psb_last_error");
#undef BIND
    d.n_sen = n_sen;
    d.n_mgau = g->n_mgau;
    d.n_feat = g->n_feat;
    d.n_density = g->n_density;
    for (f = 0; f < g->n_feat; ++f) d.featlen[f] = g->featlen[f];
    flatten_gauden(g, &mean, &var);
    d.mean = mean;
    d.var = var;
    d.det = (const float *)g->det[0][0];
    d.mixw = mixw;
    d.sen2cb = s2c;
    rc = c->model_create(&d, device, &c->model);
    if (rc == 0)
        rc = c->scorer_create(c->model, n_hist, &c->scorer);
    if (rc != 0) {
        E_ERROR("cuda_mgau: %s\n", c->last_error());
        goto fail;
    }
    ckd_free(mean); ckd_free(var); ckd_free(mixw); ckd_free(s2c); ckd_free(wide);
    (void)acmod;
    c->host = host;
    c->g = g;
    c->n_sen = n_sen;
    c->base.vt = &cuda_mgau_funcs;
    c->base.frame_idx = host->frame_idx;
    E_INFO("cuda_mgau: %s model on device %d (%d codebooks x %d streams x %d densities, %d senones)\n",
           name, device, g->n_mgau, g->n_feat, g->n_density, n_sen);
    return &c->base;
fail:
    ckd_free(mean); ckd_free(var); ckd_free(mixw); ckd_free(s2c); ckd_free(wide);
    if (c->model) c->model_free(c->model);
    if (c->dl) dlclose(c->dl);
    ckd_free(c);
    return NULL;
}

static int
cuda_mgau_frame_eval(ps_mgau_t *mg, int16 *senscr, uint8 *senone_active, int32 n_senone_active,
                     mfcc_t **feat, int32 frame, int32 compallsen)
{
    cuda_mgau_t *c = (cuda_mgau_t *)mg;
    int rc;
    /* acmod_start_utt / acmod_advance / acmod_rewind write base.frame_idx directly
     * (acmod.c:419,862,874); hand the current value to the device-side scorer. */
    c->scorer_set_frame_idx(c->scorer, mg->frame_idx);
    rc = c->scorer_frame_eval(c->scorer, senscr, senone_active, n_senone_active,
                              (const float *const *)feat, frame, compallsen);
    ++c->n_calls;
    if (rc != 0) {
        /* acmod_score ignores the return value (acmod.c:1108): log and leave senscr defined */
        E_ERROR("cuda_mgau: frame %d: %s\n", frame, c->last_error());
        memset(senscr, 0, c->n_sen * sizeof(*senscr));
        return -1;
    }
    return 0;
}

static int
cuda_mgau_transform(ps_mgau_t *mg, ps_mllr_t *mllr)
{
    cuda_mgau_t *c = (cuda_mgau_t *)mg;
    gauden_t *g = c->g;
    float *mean, *var;
    int rc;
    /* the host re-reads and adapts the Gaussians (gauden_mllr_transform, ms_gauden.c:512) ... */
    if (ps_mgau_transform(c->host, mllr) < 0)
        return -1;
    /* ... and the device copy is refreshed */
    flatten_gauden(g, &mean, &var);
    rc = c->model_update(c->model, mean, var, (const float *)g->det[0][0]);
    ckd_free(mean);
    ckd_free(var);
    if (rc != 0) E_ERROR("cuda_mgau: %s\n", c->last_error());
    return rc == 0 ? 0 : -1;
}

static void
cuda_mgau_free(ps_mgau_t *mg)
{
    cuda_mgau_t *c = (cuda_mgau_t *)mg;
    if (c == NULL) return;
    c->scorer_free(c->scorer);
    c->model_free(c->model);
    ps_mgau_free(c->host);
    dlclose(c->dl);
    ckd_free(c);
}

long
cuda_mgau_n_calls(ps_mgau_t *mg)
{
    return ((cuda_mgau_t *)mg)->n_calls;
}
