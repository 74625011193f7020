/* oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A thin, ctypes-friendly driver over the UNMODIFIED reference (cmusphinx/pocketsphinx
 * 5.1.1), compiled from the sources where they lie under /root/reference by oracle/Makefile
 * into oracle/_ref/libpsref.so.  Nothing here is shipped or linked into the product: it is
 * the checker (golden-vector generator, parity oracle, and the "reference" CPU baseline).
 *
 * It reaches the reference through the same internal headers its own unit tests use
 * (test/unit/CMakeLists.txt:60-64): acmod.h, ptm_mgau.h, s2_semi_mgau.h, ms_mgau.h, hmm.h,
 * phone_loop_search.h.  All arithmetic is the reference's; this file only moves arrays.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <pocketsphinx.h>

#include "pocketsphinx_internal.h"
#include "acmod.h"
#include "ptm_mgau.h"
#include "s2_semi_mgau.h"
#include "ms_mgau.h"
#include "hmm.h"
#include "tmat.h"
#include "bin_mdef.h"
#include "phone_loop_search.h"
#include "util/ckd_alloc.h"
#include "tied_mgau_common.h"
#include "fe/fe_internal.h"
#include "fe/fe_noise.h"
#include "feat/feat.h"

/* s2_semi_mgau.c:64-67 keeps this struct private; layout restated for history reset/dump. */
struct vqFeature_s {
    int32 score;
    int32 codeword;
};

enum { KIND_PTM = 0, KIND_SEMI = 1, KIND_MS = 2 };

typedef struct refdrv_s {
    ps_config_t *config;
    logmath_t *lmath;
    acmod_t *acmod;
    int kind;
    int sumlen;
    ps_search_t *pls;
    /* CMN state right after init: restored before every utterance so that each one is
     * featurised like the first utterance of a fresh decoder (live CMN carries over otherwise). */
    mfcc_t cmn_mean0[64], cmn_sum0[64];
    int32 cmn_nframe0;
} refdrv_t;

static double
now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static gauden_t *
drv_gauden(refdrv_t *d)
{
    switch (d->kind) {
    case KIND_PTM: return ((ptm_mgau_t *)d->acmod->mgau)->g;
    case KIND_SEMI: return ((s2_semi_mgau_t *)d->acmod->mgau)->g;
    default: return ((ms_mgau_model_t *)d->acmod->mgau)->g;
    }
}

static void
cmn_snapshot(refdrv_t *d, int restore)
{
    cmn_t *c = d->acmod->fcb->cmn_struct;
    int i;
    if (c == NULL) return;
    for (i = 0; i < c->veclen && i < 64; ++i) {
        if (restore) { c->cmn_mean[i] = d->cmn_mean0[i]; c->sum[i] = d->cmn_sum0[i]; }
        else { d->cmn_mean0[i] = c->cmn_mean[i]; d->cmn_sum0[i] = c->sum[i]; }
    }
    if (restore) c->nframe = d->cmn_nframe0; else d->cmn_nframe0 = c->nframe;
}

/* kv: newline-separated "key=value" overrides applied after feat.params. */
refdrv_t *
refdrv_open(const char *hmmdir, const char *kv)
{
    refdrv_t *d = calloc(1, sizeof(*d));
    const char *name;
    gauden_t *g;
    int i;

    err_set_loglevel(ERR_ERROR);
    d->config = ps_config_init(NULL);
    ps_config_set_str(d->config, "hmm", hmmdir);
    ps_config_set_str(d->config, "lm", NULL);
    ps_config_set_bool(d->config, "compallsen", TRUE);
    ps_config_set_int(d->config, "pl_window", 0);
    ps_config_set_str(d->config, "dither", "no");
    ps_expand_model_config(d->config);
    if (kv) {
        char *buf = strdup(kv), *save = NULL, *tok;
        for (tok = strtok_r(buf, "\n", &save); tok; tok = strtok_r(NULL, "\n", &save)) {
            char *eq = strchr(tok, '=');
            if (!eq) continue;
            *eq = 0;
            if (ps_config_set_str(d->config, tok, eq[1] ? eq + 1 : NULL) == NULL && eq[1])
                fprintf(stderr, "refdrv: could not set %s\n", tok);
        }
        free(buf);
    }
    d->lmath = logmath_init(ps_config_float(d->config, "logbase"), 0, TRUE);
    d->acmod = acmod_init(d->config, d->lmath, NULL, NULL);
    if (d->acmod == NULL) {
        fprintf(stderr, "refdrv: acmod_init failed for %s\n", hmmdir);
        free(d);
        return NULL;
    }
    acmod_set_grow(d->acmod, TRUE);
    cmn_snapshot(d, 0);
    name = d->acmod->mgau->vt->name;
    d->kind = !strcmp(name, "ptm") ? KIND_PTM : !strcmp(name, "s2_semi") ? KIND_SEMI : KIND_MS;
    g = drv_gauden(d);
    for (i = 0; i < g->n_feat; ++i)
        d->sumlen += g->featlen[i];
    return d;
}

void
refdrv_close(refdrv_t *d)
{
    if (!d) return;
    if (d->pls) ps_search_free(d->pls);
    acmod_free(d->acmod);
    logmath_free(d->lmath);
    ps_config_free(d->config);
    free(d);
}

/* dims[]: 0 kind, 1 n_sen, 2 n_mgau, 3 n_feat, 4 n_density, 5 topn, 6 sumlen, 7 n_emit_state,
 * 8 n_tmat, 9 n_sseq, 10 n_ciphone, 11 n_ci_sen, 12 mixw_4bit, 13 logadd8 size, 14 ds_ratio,
 * 15 aw, 16 ms logadd size, 17 ms logadd width, 18 ms logadd zero, 19 n_hist, 20.. featlen[] */
int
refdrv_dims(refdrv_t *d, int32 *out)
{
    gauden_t *g = drv_gauden(d);
    bin_mdef_t *m = d->acmod->mdef;
    int i;

    memset(out, 0, 32 * sizeof(*out));
    out[0] = d->kind;
    out[1] = bin_mdef_n_sen(m);
    out[2] = g->n_mgau;
    out[3] = g->n_feat;
    out[4] = g->n_density;
    out[6] = d->sumlen;
    out[7] = m->n_emit_state;
    out[8] = d->acmod->tmat->n_tmat;
    out[9] = m->n_sseq;
    out[10] = m->n_ciphone;
    out[11] = m->n_ci_sen;
    if (d->kind == KIND_PTM) {
        ptm_mgau_t *s = (ptm_mgau_t *)d->acmod->mgau;
        out[5] = s->max_topn;
        out[12] = s->mixw_cb != NULL;
        out[13] = LOGMATH_TABLE(s->lmath_8b)->table_size;
        out[14] = s->ds_ratio;
        out[19] = s->n_fast_hist;
    }
    else if (d->kind == KIND_SEMI) {
        s2_semi_mgau_t *s = (s2_semi_mgau_t *)d->acmod->mgau;
        out[5] = s->max_topn;
        out[12] = s->mixw_cb != NULL;
        out[13] = LOGMATH_TABLE(s->lmath_8b)->table_size;
        out[14] = s->ds_ratio;
        out[19] = s->n_topn_hist;
    }
    else {
        ms_mgau_model_t *s = (ms_mgau_model_t *)d->acmod->mgau;
        logadd_t *t = LOGMATH_TABLE(s->s->lmath);
        out[5] = s->topn;
        out[15] = s->s->aw;
        out[16] = t->table_size;
        out[17] = t->width;
        out[18] = logmath_get_zero(s->s->lmath);
    }
    for (i = 0; i < g->n_feat && i < 12; ++i)
        out[20 + i] = g->featlen[i];
    return 0;
}

/* Copy a named model array into out (cap bytes); returns bytes needed, or -1. */
long
refdrv_export(refdrv_t *d, const char *what, void *out, long cap)
{
    gauden_t *g = drv_gauden(d);
    bin_mdef_t *m = d->acmod->mdef;
    tmat_t *tm = d->acmod->tmat;
    long need = -1;
    int i, f, c, k;

#define EMIT(ptr, nbytes) do { need = (long)(nbytes); \
        if (out && cap >= need) memcpy(out, (ptr), need); } while (0)

    if (!strcmp(what, "mean") || !strcmp(what, "var")) {
        /* [n_mgau][n_feat][n_density][featlen[f]] floats, as precomputed by
         * gauden_dist_precompute (ms_gauden.c:264-308). */
        mfcc_t ****src = !strcmp(what, "mean") ? g->mean : g->var;
        char *o = out;
        need = (long)g->n_mgau * g->n_density * d->sumlen * sizeof(mfcc_t);
        if (out && cap >= need)
            for (i = 0; i < g->n_mgau; ++i)
                for (f = 0; f < g->n_feat; ++f) {
                    size_t nb = (size_t)g->n_density * g->featlen[f] * sizeof(mfcc_t);
                    memcpy(o, src[i][f][0], nb);
                    o += nb;
                }
    }
    else if (!strcmp(what, "det")) {
        EMIT(g->det[0][0], (long)g->n_mgau * g->n_feat * g->n_density * sizeof(mfcc_t));
    }
    else if (!strcmp(what, "mixw")) {
        uint8 ***mixw = NULL; uint8 *cb = NULL; int n_sen = bin_mdef_n_sen(m);
        if (d->kind == KIND_PTM) { mixw = ((ptm_mgau_t *)d->acmod->mgau)->mixw; cb = ((ptm_mgau_t *)d->acmod->mgau)->mixw_cb; }
        else if (d->kind == KIND_SEMI) { mixw = ((s2_semi_mgau_t *)d->acmod->mgau)->mixw; cb = ((s2_semi_mgau_t *)d->acmod->mgau)->mixw_cb; }
        if (mixw) {
            /* [n_feat][n_density][row] with row = n_sen bytes (8-bit) or (n_sen+1)/2 (4-bit). */
            long row = cb ? (n_sen + 1) / 2 : n_sen;
            char *o = out;
            need = (long)g->n_feat * g->n_density * row;
            if (out && cap >= need)
                for (f = 0; f < g->n_feat; ++f)
                    for (c = 0; c < g->n_density; ++c) {
                        memcpy(o, mixw[f][c], row);
                        o += row;
                    }
        }
        else {
            /* ms: senone_t.pdf, either [sen][feat][cw] (n_gauden>1) or [feat][cw][sen]. */
            senone_t *s = ((ms_mgau_model_t *)d->acmod->mgau)->s;
            EMIT(s->pdf[0][0], (long)s->n_sen * s->n_feat * s->n_cw);
        }
    }
    else if (!strcmp(what, "mixw_cb")) {
        uint8 *cb = NULL;
        if (d->kind == KIND_PTM) cb = ((ptm_mgau_t *)d->acmod->mgau)->mixw_cb;
        else if (d->kind == KIND_SEMI) cb = ((s2_semi_mgau_t *)d->acmod->mgau)->mixw_cb;
        if (cb) EMIT(cb, 16); else need = 0;
    }
    else if (!strcmp(what, "sen2cb")) {
        int n_sen = bin_mdef_n_sen(m);
        int32 *tmp = calloc(n_sen, sizeof(*tmp));
        for (i = 0; i < n_sen; ++i) {
            if (d->kind == KIND_PTM) tmp[i] = ((ptm_mgau_t *)d->acmod->mgau)->sen2cb[i];
            else if (d->kind == KIND_SEMI) tmp[i] = 0;
            else tmp[i] = ((ms_mgau_model_t *)d->acmod->mgau)->s->mgau[i];
        }
        EMIT(tmp, (long)n_sen * sizeof(*tmp));
        free(tmp);
    }
    else if (!strcmp(what, "logadd8")) {
        logmath_t *l8 = d->kind == KIND_PTM ? ((ptm_mgau_t *)d->acmod->mgau)->lmath_8b
            : d->kind == KIND_SEMI ? ((s2_semi_mgau_t *)d->acmod->mgau)->lmath_8b : NULL;
        if (l8) EMIT(LOGMATH_TABLE(l8)->table, LOGMATH_TABLE(l8)->table_size); else need = 0;
    }
    else if (!strcmp(what, "logadd_ms")) {
        if (d->kind == KIND_MS) {
            logadd_t *t = LOGMATH_TABLE(((ms_mgau_model_t *)d->acmod->mgau)->s->lmath);
            EMIT(t->table, (long)t->table_size * t->width);
        } else need = 0;
    }
    else if (!strcmp(what, "topn_beam")) {
        if (d->kind == KIND_SEMI) EMIT(((s2_semi_mgau_t *)d->acmod->mgau)->topn_beam, g->n_feat);
        else need = 0;
    }
    else if (!strcmp(what, "tp")) {
        /* uint8 [n_tmat][n_state][n_state+1] (tmat.h:57-63); rows are contiguous. */
        EMIT(tm->tp[0][0], (long)tm->n_tmat * tm->n_state * (tm->n_state + 1));
    }
    else if (!strcmp(what, "sseq")) {
        uint16 *tmp = calloc((size_t)m->n_sseq * m->n_emit_state, sizeof(*tmp));
        for (i = 0; i < m->n_sseq; ++i)
            for (k = 0; k < m->n_emit_state; ++k)
                tmp[i * m->n_emit_state + k] = m->sseq[i][k];
        EMIT(tmp, (long)m->n_sseq * m->n_emit_state * sizeof(*tmp));
        free(tmp);
    }
    else if (!strcmp(what, "phone_ssid") || !strcmp(what, "phone_tmat")) {
        /* per phone (CI first, then CD): ssid / tmat (bin_mdef.h:159-160). */
        int32 *tmp = calloc(m->n_phone, sizeof(*tmp));
        for (i = 0; i < m->n_phone; ++i)
            tmp[i] = !strcmp(what, "phone_ssid") ? bin_mdef_pid2ssid(m, i) : bin_mdef_pid2tmatid(m, i);
        EMIT(tmp, (long)m->n_phone * sizeof(*tmp));
        free(tmp);
    }
#undef EMIT
    return need;
}

/* Front-end parameters and tables exactly as fe_init / fe_build_melfilters / fe_compute_melcosine
 * left them (fe_internal.h:100-180).  out: int32[16] then float[4]:
 *  [0] frame_size [1] frame_shift [2] fft_size [3] fft_order [4] num_filters [5] num_cepstra
 *  [6] remove_dc [7] remove_noise [8] transform [9] lifter_val [10] log_spec [11] dither
 *  [12] n_filt_coeffs [13] feat window_size [14] cmn type [15] feat cepsize;
 *  f[0] pre_emphasis_alpha f[1] sqrt_inv_n f[2] sqrt_inv_2n f[3] sampling_rate */
int
refdrv_fe_info(refdrv_t *d, int32 *out, float *fout)
{
    fe_t *fe = d->acmod->fe;
    melfb_t *mf = fe->mel_fb;
    int i, n = 0;
    for (i = 0; i < mf->num_filters; ++i) n += mf->filt_width[i];
    out[0] = fe->frame_size; out[1] = fe->frame_shift; out[2] = fe->fft_size; out[3] = fe->fft_order;
    out[4] = mf->num_filters; out[5] = fe->num_cepstra; out[6] = fe->remove_dc; out[7] = fe->noise_stats != NULL;
    out[8] = fe->transform; out[9] = mf->lifter_val; out[10] = fe->log_spec; out[11] = fe->dither;
    out[12] = n; out[13] = feat_window_size(d->acmod->fcb); out[14] = d->acmod->fcb->cmn; out[15] = feat_cepsize(d->acmod->fcb);
    fout[0] = fe->pre_emphasis_alpha; fout[1] = mf->sqrt_inv_n; fout[2] = mf->sqrt_inv_2n; fout[3] = fe->sampling_rate;
    return 0;
}

long
refdrv_fe_export(refdrv_t *d, const char *what, void *out, long cap)
{
    fe_t *fe = d->acmod->fe;
    melfb_t *mf = fe->mel_fb;
    long need = -1;
    int i, n = 0;
#define EMIT(ptr, nbytes) do { need = (long)(nbytes); \
        if (out && cap >= need) memcpy(out, (ptr), need); } while (0)
    for (i = 0; i < mf->num_filters; ++i) n += mf->filt_width[i];
    if (!strcmp(what, "hamming")) EMIT(fe->hamming_window, (long)(fe->frame_size / 2) * sizeof(window_t));
    else if (!strcmp(what, "ccc")) EMIT(fe->ccc, (long)(fe->fft_size / 4) * sizeof(frame_t));
    else if (!strcmp(what, "sss")) EMIT(fe->sss, (long)(fe->fft_size / 4) * sizeof(frame_t));
    else if (!strcmp(what, "spec_start")) EMIT(mf->spec_start, (long)mf->num_filters * sizeof(int16));
    else if (!strcmp(what, "filt_start")) EMIT(mf->filt_start, (long)mf->num_filters * sizeof(int16));
    else if (!strcmp(what, "filt_width")) EMIT(mf->filt_width, (long)mf->num_filters * sizeof(int16));
    else if (!strcmp(what, "filt_coeffs")) EMIT(mf->filt_coeffs, (long)n * sizeof(mfcc_t));
    else if (!strcmp(what, "mel_cosine")) EMIT(mf->mel_cosine[0], (long)mf->num_cepstra * mf->num_filters * sizeof(mfcc_t));
    else if (!strcmp(what, "lifter")) { if (mf->lifter_val) EMIT(mf->lifter, (long)mf->num_cepstra * sizeof(mfcc_t)); else need = 0; }
#undef EMIT
    return need;
}

/* PCM -> cepstra (before CMN) through fe_process_frames + fe_end_utt, noise tracker reset first
 * (ps_start_stream, pocketsphinx.c:1073-1083): what a fresh stream produces.  out [T][num_cepstra]. */
int
refdrv_mfcc(refdrv_t *d, const int16 *pcm, long n_samples, float *out, int max_frames)
{
    fe_t *fe = d->acmod->fe;
    const int16 *p = pcm;
    size_t n = n_samples;
    int32 nfr = 0, T, t, last = 0;
    mfcc_t **buf;

    fe_reset_noisestats(fe->noise_stats);
    fe_start_utt(fe);
    fe_process_frames(fe, NULL, &n, NULL, &nfr);
    T = nfr + 1;
    buf = (mfcc_t **)ckd_calloc_2d(T, fe->num_cepstra, sizeof(mfcc_t));
    nfr = T;
    p = pcm; n = n_samples;
    if (n_samples > 0 && T > 1) fe_process_frames(fe, &p, &n, buf, &nfr);
    else {
        /* fewer samples than one frame: fe_process_frames only buffers them */
        int32 z = 1;
        fe_process_frames(fe, &p, &n, buf, &z);
        nfr = 0;
    }
    fe_end_utt(fe, buf[nfr], &last);
    nfr += last;
    for (t = 0; t < nfr && t < max_frames; ++t)
        memcpy(out + (size_t)t * fe->num_cepstra, buf[t], fe->num_cepstra * sizeof(mfcc_t));
    ckd_free_2d(buf);
    return nfr;
}

void
refdrv_fe_reset(refdrv_t *d)
{
    fe_reset_noisestats(d->acmod->fe->noise_stats);
}

/* PCM -> dynamic features through the reference fe/ + feat/ (full-utterance mode, as
 * ps_decode_raw does).  out is [T][sumlen] floats.  Returns the number of frames. */
int
refdrv_featurize(refdrv_t *d, const int16 *pcm, long n_samples, float *out, int max_frames)
{
    acmod_t *a = d->acmod;
    const int16 *p = pcm;
    size_t n = n_samples;
    int T, t;

    cmn_snapshot(d, 1);
    acmod_start_utt(a);
    acmod_process_raw(a, &p, &n, TRUE);
    acmod_end_utt(a);
    T = a->n_feat_frame;
    if (T > max_frames) T = max_frames;
    for (t = 0; t < T; ++t)
        memcpy(out + (size_t)t * d->sumlen, a->feat_buf[t][0], d->sumlen * sizeof(float));
    return a->n_feat_frame;
}

/* Restore the top-N history to its post-init state (ptm_mgau.c:777-803, s2_semi_mgau.c:1319-1327)
 * in place, so every utterance starts like a fresh decoder. */
void
refdrv_reset(refdrv_t *d)
{
    ps_mgau_t *mg = d->acmod->mgau;
    int i, j, k, m;

    mg->frame_idx = 0;
    if (d->kind == KIND_PTM) {
        ptm_mgau_t *s = (ptm_mgau_t *)mg;
        for (i = 0; i < s->n_fast_hist; ++i) {
            for (j = 0; j < s->g->n_mgau; ++j)
                for (k = 0; k < s->g->n_feat; ++k)
                    for (m = 0; m < s->max_topn; ++m) {
                        s->hist[i].topn[j][k][m].cw = m;
                        s->hist[i].topn[j][k][m].score = WORST_DIST;
                    }
            bitvec_set_all(s->hist[i].mgau_active, s->g->n_mgau);
        }
    }
    else if (d->kind == KIND_SEMI) {
        s2_semi_mgau_t *s = (s2_semi_mgau_t *)mg;
        for (i = 0; i < s->n_topn_hist; ++i)
            for (j = 0; j < s->g->n_feat; ++j) {
                for (k = 0; k < s->max_topn; ++k) {
                    s->topn_hist[i][j][k].score = WORST_DIST;
                    s->topn_hist[i][j][k].codeword = k;
                }
                s->topn_hist_n[i][j] = 0;
            }
    }
}

static void
feat_ptrs(refdrv_t *d, const float *row, mfcc_t **ptrs)
{
    gauden_t *g = drv_gauden(d);
    int f, off = 0;
    for (f = 0; f < g->n_feat; ++f) {
        ptrs[f] = (mfcc_t *)row + off;
        off += g->featlen[f];
    }
}

/* Score T frames of features with the reference back-end's own frame_eval, all senones
 * (compallsen), frame_idx advanced like acmod_advance (acmod.c:868-877).
 * topn_out (optional): PTM [T][n_mgau][n_feat][topn][2] / SEMI [T][n_feat][topn][2] int32
 * {cw, normalised score} after each frame. */
int
refdrv_score(refdrv_t *d, const float *feats, int T, int16 *senscr, int reset, int32 *topn_out)
{
    ps_mgau_t *mg = d->acmod->mgau;
    int n_sen = bin_mdef_n_sen(d->acmod->mdef);
    mfcc_t *ptrs[16];
    int t, base;

    if (reset) refdrv_reset(d);
    base = mg->frame_idx;
    for (t = 0; t < T; ++t) {
        feat_ptrs(d, feats + (size_t)t * d->sumlen, ptrs);
        ps_mgau_frame_eval(mg, senscr + (size_t)t * n_sen, NULL, 0, ptrs, base + t, TRUE);
        mg->frame_idx = base + t + 1;
        if (topn_out && d->kind == KIND_PTM) {
            ptm_mgau_t *s = (ptm_mgau_t *)mg;
            size_t n = (size_t)s->g->n_mgau * s->g->n_feat * s->max_topn;
            memcpy(topn_out + t * n * 2, s->f->topn[0][0], n * sizeof(ptm_topn_t));
        }
        else if (topn_out && d->kind == KIND_SEMI) {
            s2_semi_mgau_t *s = (s2_semi_mgau_t *)mg;
            size_t n = (size_t)s->g->n_feat * s->max_topn;
            int f, k;
            for (f = 0; f < s->g->n_feat; ++f)
                for (k = 0; k < s->max_topn; ++k) {
                    topn_out[(t * n + f * s->max_topn + k) * 2] = s->f[f][k].codeword;
                    topn_out[(t * n + f * s->max_topn + k) * 2 + 1] = s->f[f][k].score;
                }
        }
    }
    return 0;
}

/* Same, but with per-frame active-senone flags (compallsen = no): the flags go through the
 * reference's own bitvec -> delta-list coder (acmod_flags2list, acmod.c:1224-1275).
 * flags: [T][n_sen] bytes.  n_active_out/list_out (optional): the delta list per frame. */
int
refdrv_score_active(refdrv_t *d, const float *feats, int T, const uint8 *flags,
                    int16 *senscr, int reset, int32 *n_active_out, uint8 *list_out)
{
    acmod_t *a = d->acmod;
    ps_mgau_t *mg = a->mgau;
    int n_sen = bin_mdef_n_sen(a->mdef);
    mfcc_t *ptrs[16];
    int t, s, base, n;
    uint8 save = a->compallsen;

    if (reset) refdrv_reset(d);
    base = mg->frame_idx;
    a->compallsen = FALSE;
    for (t = 0; t < T; ++t) {
        acmod_clear_active(a);
        for (s = 0; s < n_sen; ++s)
            if (flags[(size_t)t * n_sen + s])
                bitvec_set(a->senone_active_vec, s);
        n = acmod_flags2list(a);
        if (n_active_out) n_active_out[t] = n;
        if (list_out) memcpy(list_out + (size_t)t * n_sen, a->senone_active, n);
        feat_ptrs(d, feats + (size_t)t * d->sumlen, ptrs);
        ps_mgau_frame_eval(mg, senscr + (size_t)t * n_sen, a->senone_active, n, ptrs, base + t, FALSE);
        mg->frame_idx = base + t + 1;
    }
    a->compallsen = save;
    return 0;
}

/* CPU baseline: wall-clock seconds for reps passes of refdrv_score over the same features. */
double
refdrv_time_score(refdrv_t *d, const float *feats, int T, int reps)
{
    int n_sen = bin_mdef_n_sen(d->acmod->mdef);
    int16 *scr = malloc((size_t)T * n_sen * sizeof(*scr));
    double t0, t1;
    int r;

    refdrv_score(d, feats, T < 16 ? T : 16, scr, 1, NULL); /* warm caches */
    t0 = now_s();
    for (r = 0; r < reps; ++r)
        refdrv_score(d, feats, T, scr, 1, NULL);
    t1 = now_s();
    free(scr);
    return t1 - t0;
}

/* ------------------------------------------------------------------------------------- */
/* HMM evaluation through the reference's hmm.c                                           */

typedef struct refhmmctx_s {
    hmm_context_t *ctx;
    uint8 ***tp;
    uint16 **sseq;
    int n_tmat, n_sseq, n_emit;
} refhmmctx_t;

/* tp_flat: uint8 [n_tmat][n_emit][n_emit+1]; sseq_flat: uint16 [n_sseq][n_emit]. */
refhmmctx_t *
refdrv_hmmctx_new(int n_emit, const uint8 *tp_flat, int n_tmat, const uint16 *sseq_flat, int n_sseq)
{
    refhmmctx_t *c = calloc(1, sizeof(*c));
    c->n_emit = n_emit;
    c->n_tmat = n_tmat;
    c->n_sseq = n_sseq;
    c->tp = (uint8 ***)ckd_calloc_3d(n_tmat, n_emit, n_emit + 1, 1);
    memcpy(c->tp[0][0], tp_flat, (size_t)n_tmat * n_emit * (n_emit + 1));
    c->sseq = (uint16 **)ckd_calloc_2d(n_sseq, n_emit, sizeof(uint16));
    memcpy(c->sseq[0], sseq_flat, (size_t)n_sseq * n_emit * sizeof(uint16));
    c->ctx = hmm_context_init(n_emit, c->tp, NULL, c->sseq);
    return c;
}

void
refdrv_hmmctx_free(refhmmctx_t *c)
{
    if (!c) return;
    hmm_context_free(c->ctx);
    ckd_free_3d(c->tp);
    ckd_free_2d(c->sseq);
    free(c);
}

int refdrv_sizeof_hmm(void) { return (int)sizeof(hmm_t); }

/* hmms: array of n real hmm_t records (88 bytes each; the ctx field is overwritten here).
 * Calls hmm_vit_eval (hmm.c:787) on each and returns max bestscore (WORST_SCORE if n==0),
 * like evaluate_hmms (phone_loop_search.c:202-221). */
int32
refdrv_hmm_vit_eval(refhmmctx_t *c, void *hmms, int n, const int16 *senscr)
{
    hmm_t *h = hmms;
    int32 best = WORST_SCORE;
    int i;
    hmm_context_set_senscore(c->ctx, senscr);
    for (i = 0; i < n; ++i) {
        int32 s;
        h[i].ctx = c->ctx;
        s = hmm_vit_eval(&h[i]);
        if (s BETTER_THAN best) best = s;
    }
    return best;
}

/* The evaluate_channels loop (ngram_search_fwdtree.c:702-715) over a fixed active set for T frames:
 * frame t scores the n hmm_t against row t of senscr ([T][n_sen] int16) with the reference's
 * hmm_vit_eval and records the best score -- the CPU side of bench.py's search-scale Viterbi. */
void
refdrv_hmm_sweep(refhmmctx_t *c, void *hmms, int n, const int16 *senscr, int n_sen, int T, int32 *best_out)
{
    int t;
    for (t = 0; t < T; ++t)
        best_out[t] = refdrv_hmm_vit_eval(c, hmms, n, senscr + (size_t)t * n_sen);
}

/* hmm_init (hmm.c:85-105) on caller-provided storage. */
void
refdrv_hmm_init(refhmmctx_t *c, void *hmms, int n, const int32 *mpx, const int32 *ssid, const int32 *tmatid)
{
    hmm_t *h = hmms;
    int i;
    for (i = 0; i < n; ++i)
        hmm_init(c->ctx, &h[i], mpx[i], ssid[i], tmatid[i]);
}

void refdrv_hmm_enter(void *hmms, int i, int32 score, int32 histid, int frame) { hmm_enter((hmm_t *)hmms + i, score, histid, frame); }
void refdrv_hmm_clear(void *hmms, int i) { hmm_clear((hmm_t *)hmms + i); }
void refdrv_hmm_clear_scores(void *hmms, int i) { hmm_clear_scores((hmm_t *)hmms + i); }
void refdrv_hmm_normalize(void *hmms, int i, int32 best) { hmm_normalize((hmm_t *)hmms + i, best); }

double
refdrv_time_hmm_vit_eval(refhmmctx_t *c, void *hmms, int n, const int16 *senscr, int reps)
{
    void *copy = malloc((size_t)n * sizeof(hmm_t));
    double t0, t = 0;
    int r;
    for (r = 0; r < reps; ++r) {
        memcpy(copy, hmms, (size_t)n * sizeof(hmm_t));
        t0 = now_s();
        refdrv_hmm_vit_eval(c, copy, n, senscr);
        t += now_s() - t0;
    }
    free(copy);
    return t;
}

/* ------------------------------------------------------------------------------------- */
/* The reference's own phone loop (phone_loop_search.c) run over PCM; per frame dumps
 * every phone HMM (as 88-byte hmm_t), best_score and penalties.                          */

int
refdrv_phoneloop_run(refdrv_t *d, const int16 *pcm, long n_samples, const char *kv,
                     int max_frames, void *hmm_out, int32 *best_out, int32 *pen_out,
                     int16 *senscr_out, float *feat_out)
{
    acmod_t *a = d->acmod;
    phone_loop_search_t *pls;
    const int16 *p = pcm;
    size_t n = n_samples;
    int T, t, np, n_sen = bin_mdef_n_sen(a->mdef);

    if (kv) {
        char *buf = strdup(kv), *save = NULL, *tok;
        for (tok = strtok_r(buf, "\n", &save); tok; tok = strtok_r(NULL, "\n", &save)) {
            char *eq = strchr(tok, '=');
            if (!eq) continue;
            *eq = 0;
            ps_config_set_str(d->config, tok, eq + 1);
        }
        free(buf);
    }
    if (d->pls) ps_search_free(d->pls);
    d->pls = phone_loop_search_init(d->config, a, NULL);
    pls = (phone_loop_search_t *)d->pls;
    np = pls->n_phones;

    refdrv_reset(d);
    cmn_snapshot(d, 1);
    acmod_start_utt(a);
    acmod_process_raw(a, &p, &n, TRUE);
    acmod_end_utt(a);
    T = a->n_feat_frame;
    ps_search_start(d->pls);
    for (t = 0; t < T && t < max_frames; ++t) {
        int fi = t;
        if (feat_out)
            memcpy(feat_out + (size_t)t * d->sumlen, a->feat_buf[t][0], d->sumlen * sizeof(float));
        ps_search_step(d->pls, t);
        if (senscr_out)
            memcpy(senscr_out + (size_t)t * n_sen, acmod_score(a, &fi), n_sen * sizeof(int16));
        if (hmm_out)
            memcpy((char *)hmm_out + (size_t)t * np * sizeof(hmm_t), pls->hmms, np * sizeof(hmm_t));
        if (best_out) best_out[t] = pls->best_score;
        if (pen_out) memcpy(pen_out + (size_t)t * np, pls->penalties, np * sizeof(int32));
        acmod_advance(a);
    }
    ps_search_finish(d->pls);
    return T;
}

int
refdrv_phoneloop_params(refdrv_t *d, int32 *out)
{
    phone_loop_search_t *pls = (phone_loop_search_t *)d->pls;
    if (!pls) return -1;
    out[0] = pls->n_phones;
    out[1] = pls->beam;
    out[2] = pls->pbeam;
    out[3] = pls->pip;
    out[4] = pls->window;
    memcpy(out + 6, &pls->penalty_weight, sizeof(double)); /* out[6..7] = float64 */
    return 0;
}

/* ------------------------------------------------------------------------------------- */
/* Full decoder (ps_init / ps_process_raw / ps_get_hyp) with the GMM back-end optionally
 * replaced by the CUDA one through integration/ps_mgau_cuda.c: the drop-in test.          */

ps_mgau_t *cuda_mgau_wrap(acmod_t *acmod, ps_mgau_t *host, const char *libpath, int device);
long cuda_mgau_n_calls(ps_mgau_t *mg);

/* Returns the number of frames (<0 on error).  hyp/seg are NUL-terminated text; seg has one
 * "word start end ascr lscr" line per segment.  stats[0] = hypothesis score, [1] = number of
 * frame_eval calls served by the CUDA back-end (0 on the host path), [2] = n_sen, [3] = wall
 * clock in microseconds of the utterance alone (start_utt .. end_utt, second of two passes). */
int
refdrv_decode(const char *hmmdir, const char *lm, const char *dict, const char *kv,
              const int16 *pcm, long n_samples, int use_cuda, const char *libpath,
              char *hyp, int hyp_cap, char *seg, int seg_cap, int32 *stats)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    const char *h;
    int32 score = 0;
    ps_seg_t *it;
    int n, nfr;
    long calls = 0;

    err_set_loglevel(ERR_ERROR);
    config = ps_config_init(NULL);
    ps_config_set_str(config, "hmm", hmmdir);
    ps_config_set_str(config, "lm", lm);
    ps_config_set_str(config, "dict", dict);
    ps_config_set_str(config, "dither", "no");
    if (kv) {
        char *buf = strdup(kv), *save = NULL, *tok;
        for (tok = strtok_r(buf, "\n", &save); tok; tok = strtok_r(NULL, "\n", &save)) {
            char *eq = strchr(tok, '=');
            if (!eq) continue;
            *eq = 0;
            ps_config_set_str(config, tok, eq + 1);
        }
        free(buf);
    }
    ps = ps_init(config);
    if (ps == NULL) {
        ps_config_free(config);
        return -1;
    }
    if (use_cuda) {
        ps_mgau_t *g = cuda_mgau_wrap(ps->acmod, ps->acmod->mgau, libpath, 0);
        if (g == NULL) {
            ps_free(ps);
            ps_config_free(config);
            return -2;
        }
        ps->acmod->mgau = g;
    }
    {
        /* stats[3] != 0 on entry: decode the utterance twice and time the second pass (warm caches,
         * warm device; live CMN then starts from the first pass's estimate on both arms alike) */
        double t0;
        if (stats && stats[3]) {
            ps_start_utt(ps);
            ps_process_raw(ps, pcm, n_samples, FALSE, TRUE);
            ps_end_utt(ps);
        }
        t0 = now_s();
        ps_start_utt(ps);
        ps_process_raw(ps, pcm, n_samples, FALSE, TRUE);
        ps_end_utt(ps);
        if (stats) stats[3] = (int32)((now_s() - t0) * 1e6);
    }
    nfr = ps_get_n_frames(ps);
    h = ps_get_hyp(ps, &score);
    snprintf(hyp, hyp_cap, "%s", h ? h : "");
    n = 0;
    seg[0] = 0;
    for (it = ps_seg_iter(ps); it; it = ps_seg_next(it)) {
        int sf, ef;
        int32 ascr, lscr, lback;
        ps_seg_frames(it, &sf, &ef);
        ps_seg_prob(it, &ascr, &lscr, &lback);
        n += snprintf(seg + n, n < seg_cap ? seg_cap - n : 0, "%s %d %d %d %d\n", ps_seg_word(it), sf, ef, ascr, lscr);
        if (n >= seg_cap) break;
    }
    if (use_cuda) calls = cuda_mgau_n_calls(ps->acmod->mgau);
    if (stats) {
        stats[0] = score;
        stats[1] = (int32)calls;
        stats[2] = bin_mdef_n_sen(ps->acmod->mdef);
    }
    ps_free(ps);
    ps_config_free(config);
    return nfr;
}


/* ------------------------------------------------------------------------------------- */
/* Senone-dump interoperability: decode from a .sen file (ps_decode_senscr), and have the
 * reference itself write one (acmod_set_senfh) while decoding PCM with -compallsen yes.    */
int
refdrv_decode_senscr(const char *hmmdir, const char *lm, const char *dict, const char *kv, const char *senfile,
                     const int16 *pcm, long n_samples, const char *senout,
                     char *hyp, int hyp_cap, char *seg, int seg_cap, int32 *stats)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    const char *h;
    int32 score = 0;
    ps_seg_t *it;
    FILE *fh = NULL;
    int n = 0, nfr;

    err_set_loglevel(ERR_ERROR);
    config = ps_config_init(NULL);
    ps_config_set_str(config, "hmm", hmmdir);
    ps_config_set_str(config, "lm", lm);
    ps_config_set_str(config, "dict", dict);
    ps_config_set_str(config, "dither", "no");
    ps_config_set_bool(config, "compallsen", TRUE);
    if (kv) {
        char *buf = strdup(kv), *save = NULL, *tok;
        for (tok = strtok_r(buf, "\n", &save); tok; tok = strtok_r(NULL, "\n", &save)) {
            char *eq = strchr(tok, '=');
            if (!eq) continue;
            *eq = 0;
            ps_config_set_str(config, tok, eq + 1);
        }
        free(buf);
    }
    ps = ps_init(config);
    if (ps == NULL) { ps_config_free(config); return -1; }
    if (senfile) {
        fh = fopen(senfile, "rb");
        if (fh == NULL) { ps_free(ps); ps_config_free(config); return -3; }
        nfr = ps_decode_senscr(ps, fh);
        fclose(fh);
    }
    else {
        if (senout) {
            fh = fopen(senout, "wb");
            acmod_set_senfh(ps->acmod, fh);
        }
        ps_start_utt(ps);
        ps_process_raw(ps, pcm, n_samples, FALSE, TRUE);
        ps_end_utt(ps);
        nfr = ps_get_n_frames(ps);
        /* acmod_end_utt closed the dump file (acmod.c:460-463) */
    }
    h = ps_get_hyp(ps, &score);
    snprintf(hyp, hyp_cap, "%s", h ? h : "");
    seg[0] = 0;
    for (it = ps_seg_iter(ps); it; it = ps_seg_next(it)) {
        int sf, ef;
        int32 ascr, lscr, lback;
        ps_seg_frames(it, &sf, &ef);
        ps_seg_prob(it, &ascr, &lscr, &lback);
        n += snprintf(seg + n, n < seg_cap ? seg_cap - n : 0, "%s %d %d %d %d\n", ps_seg_word(it), sf, ef, ascr, lscr);
        if (n >= seg_cap) break;
    }
    if (stats) { stats[0] = score; stats[1] = 0; stats[2] = bin_mdef_n_sen(ps->acmod->mdef); }
    ps_free(ps);
    ps_config_free(config);
    return nfr;
}

/* Forced alignment through the reference's own state_align_search (state_align_search.c) on one
 * utterance: `words` (dictionary words separated by blanks, e.g. "<s> go forward ten meters </s>")
 * -> ps_alignment_t (ps_alignment_add_word with no timing, ps_alignment_populate) ->
 * ps_set_alignment -> ps_start_utt / ps_process_raw(full) / ps_end_utt.  Outputs the phone
 * sequence (ssid, tmatid) and, per emitting state, start / duration / score of the resulting
 * alignment.  info: [0] frames, [1] n_phones, [2] n_states, [3] n_emit_state.  Returns 0, or <0. */
#include "ps_alignment_internal.h"
int
refdrv_align(const char *hmmdir, const char *dict, const char *kv, const char *words,
             const int16 *pcm, long n_samples, int32 *ph_ssid, int32 *ph_tmat, int cap_ph,
             int32 *st_start, int32 *st_dur, int32 *st_score, int cap_st, int32 *info)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    ps_alignment_t *al;
    ps_alignment_iter_t *it;
    char *buf, *save = NULL, *tok;
    int i, rc = 0;

    err_set_loglevel(ERR_ERROR);
    config = ps_config_init(NULL);
    ps_config_set_str(config, "hmm", hmmdir);
    ps_config_set_str(config, "dict", dict);
    ps_config_set_str(config, "dither", "no");
    ps_config_set_str(config, "compallsen", "yes");
    ps_config_set_str(config, "pl_window", "0");
    ps_config_set_str(config, "lm", NULL);
    if (kv) {
        char *b2 = strdup(kv), *s2 = NULL, *t2;
        for (t2 = strtok_r(b2, "\n", &s2); t2; t2 = strtok_r(NULL, "\n", &s2)) {
            char *eq = strchr(t2, '=');
            if (!eq) continue;
            *eq = 0;
            ps_config_set_str(config, t2, eq + 1);
        }
        free(b2);
    }
    ps = ps_init(config);
    if (ps == NULL) { ps_config_free(config); return -1; }
    al = ps_alignment_init(ps->d2p);
    buf = strdup(words);
    for (tok = strtok_r(buf, " \t\n", &save); tok; tok = strtok_r(NULL, " \t\n", &save)) {
        int32 wid = dict_wordid(ps->dict, tok);
        if (wid == BAD_S3WID) { rc = -2; break; }
        ps_alignment_add_word(al, wid, 0, 0);
    }
    free(buf);
    if (rc == 0 && ps_alignment_populate(al) < 0) rc = -3;
    if (rc == 0 && ps_set_alignment(ps, al) < 0) rc = -4;
    if (rc == 0) {
        ps_start_utt(ps);
        ps_process_raw(ps, pcm, n_samples, FALSE, TRUE);
        if (ps_end_utt(ps) < 0) rc = -5;
        info[0] = ps_get_n_frames(ps);
        info[1] = ps_alignment_n_phones(al);
        info[2] = ps_alignment_n_states(al);
        info[3] = bin_mdef_n_emit_state(ps->acmod->mdef);
        for (i = 0, it = ps_alignment_phones(al); it; it = ps_alignment_iter_next(it), ++i) {
            ps_alignment_entry_t *e = ps_alignment_iter_get(it);
            if (i < cap_ph) { ph_ssid[i] = e->id.pid.ssid; ph_tmat[i] = e->id.pid.tmatid; }
        }
        for (i = 0, it = ps_alignment_states(al); it; it = ps_alignment_iter_next(it), ++i) {
            ps_alignment_entry_t *e = ps_alignment_iter_get(it);
            if (i < cap_st) { st_start[i] = e->start; st_dur[i] = e->duration; st_score[i] = e->score; }
        }
    }
    ps_alignment_free(al);
    ps_free(ps);
    ps_config_free(config);
    return rc;
}

/* Keyword spotting through the reference's own kws_search (kws_search.c) on one utterance:
 * `keyfile` is a kws list ("phrase /threshold/" per line) or NULL with a single `keyphrase`.
 * Exports the search's configuration in list order -- phone loop (ssid, tmatid) [n_pl], the
 * keyphrases' HMM chains kp_off[n_kp+1] / (ssid, tmatid) / threshold, beam and plp -- and the final
 * detection list (kp index, sf, ef, prob, ascr) in list order.  info: [0] frames [1] n_pl [2] n_kp
 * [3] total kp hmms [4] beam [5] plp [6] n_detections.  Returns 0 or <0. */
#include "kws_search.h"
int
refdrv_kws(const char *hmmdir, const char *dict, const char *kv, const char *keyphrase, const char *keyfile,
           const int16 *pcm, long n_samples, int32 *pl_ssid, int32 *pl_tmat, int32 *kp_off, int32 *kp_thresh,
           int32 *kp_ssid, int32 *kp_tmat, int cap, int32 *det, int cap_det, int32 *info)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    kws_search_t *kwss;
    gnode_t *gn;
    int i, k, n;

    err_set_loglevel(ERR_ERROR);
    config = ps_config_init(NULL);
    ps_config_set_str(config, "hmm", hmmdir);
    ps_config_set_str(config, "dict", dict);
    ps_config_set_str(config, "dither", "no");
    ps_config_set_str(config, "compallsen", "yes");
    ps_config_set_str(config, "pl_window", "0");
    ps_config_set_str(config, "lm", NULL);
    if (keyfile) ps_config_set_str(config, "kws", keyfile);
    else ps_config_set_str(config, "keyphrase", keyphrase);
    if (kv) {
        char *b2 = strdup(kv), *s2 = NULL, *t2;
        for (t2 = strtok_r(b2, "\n", &s2); t2; t2 = strtok_r(NULL, "\n", &s2)) {
            char *eq = strchr(t2, '=');
            if (!eq) continue;
            *eq = 0;
            ps_config_set_str(config, t2, eq + 1);
        }
        free(b2);
    }
    ps = ps_init(config);
    if (ps == NULL) { ps_config_free(config); return -1; }
    if (ps->search == NULL || strcmp(ps_search_type(ps->search), PS_SEARCH_TYPE_KWS) != 0) {
        ps_free(ps); ps_config_free(config);
        return -2;
    }
    kwss = (kws_search_t *)ps->search;
    ps_start_utt(ps);
    ps_process_raw(ps, pcm, n_samples, FALSE, TRUE);
    ps_end_utt(ps);
    info[0] = ps_get_n_frames(ps);
    info[1] = kwss->n_pl;
    info[4] = kwss->beam;
    info[5] = kwss->plp;
    for (i = 0; i < kwss->n_pl && i < cap; ++i) {
        pl_ssid[i] = hmm_nonmpx_ssid(&kwss->pl_hmms[i]);
        pl_tmat[i] = kwss->pl_hmms[i].tmatid;
    }
    k = 0; n = 0;
    kp_off[0] = 0;
    for (gn = kwss->keyphrases; gn; gn = gnode_next(gn)) {
        kws_keyphrase_t *kp = gnode_ptr(gn);
        for (i = 0; i < kp->n_hmms; ++i, ++n)
            if (n < cap) { kp_ssid[n] = hmm_nonmpx_ssid(&kp->hmms[i]); kp_tmat[n] = kp->hmms[i].tmatid; }
        kp_thresh[k] = kp->threshold;
        kp_off[++k] = n;
    }
    info[2] = k;
    info[3] = n;
    n = 0;
    for (gn = kwss->detections->detect_list; gn; gn = gnode_next(gn)) {
        kws_detection_t *d = gnode_ptr(gn);
        gnode_t *g2;
        int idx = 0, which = -1;
        for (g2 = kwss->keyphrases; g2; g2 = gnode_next(g2), ++idx)
            if (strcmp(((kws_keyphrase_t *)gnode_ptr(g2))->word, d->keyphrase) == 0) which = idx;
        if (n < cap_det) {
            det[n * 5 + 0] = which; det[n * 5 + 1] = d->sf; det[n * 5 + 2] = d->ef;
            det[n * 5 + 3] = d->prob; det[n * 5 + 4] = d->ascr;
        }
        ++n;
    }
    info[6] = n;
    ps_free(ps);
    ps_config_free(config);
    return 0;
}

/* Phone decoding through the reference's own allphone_search (allphone_search.c) without a phone
 * LM (unconstrained loop, insertion penalty only) on one utterance.  Exports the search graph in
 * the order phmm_eval_all / phmm_exit walk it (ci-major, list order): per node (ci, ssid, tmatid),
 * successor lists in CSR form, the start node, beam / pbeam / inspen, and the resulting phone
 * segmentation (ci, sf, ef, score, tscore).  info: [0] frames [1] n_nodes [2] n_links [3] start
 * [4] beam [5] pbeam [6] inspen [7] n_segments [8] n_history.  Returns 0 or <0. */
#include "allphone_search.h"
/* lm_tables (may be NULL; used when kv sets "allphone" to a phone LM): int32 [n_ci*n_ci] bigram
 * then [n_ci*n_ci*n_ci] trigram scores >> SENSCR_SHIFT, tabulated through the search's own LM
 * object with the argument positions of phmm_exit / phmm_trans (allphone_search.c:420-441,
 * 497-513): bg[a][b] = ngram_bg_score(lm, wid[a], wid[b]), tg[a][b][c] = ngram_tg_score(lm,
 * wid[a], wid[b], wid[c]).  info[9] = 1 if an LM is in use, info[10] = n_ci. */
int
refdrv_allphone_lm(const char *hmmdir, const char *kv, const int16 *pcm, long n_samples,
                   int32 *node_ci, int32 *node_ssid, int32 *node_tmat, int32 *succ_off, int cap_nodes,
                   int32 *succ, int cap_links, int32 *segs, int cap_segs, int32 *info, int32 *lm_tables);

int
refdrv_allphone(const char *hmmdir, const char *kv, const int16 *pcm, long n_samples,
                int32 *node_ci, int32 *node_ssid, int32 *node_tmat, int32 *succ_off, int cap_nodes,
                int32 *succ, int cap_links, int32 *segs, int cap_segs, int32 *info)
{
    return refdrv_allphone_lm(hmmdir, kv, pcm, n_samples, node_ci, node_ssid, node_tmat, succ_off, cap_nodes, succ,
                              cap_links, segs, cap_segs, info, NULL);
}

int
refdrv_allphone_lm(const char *hmmdir, const char *kv, const int16 *pcm, long n_samples,
                   int32 *node_ci, int32 *node_ssid, int32 *node_tmat, int32 *succ_off, int cap_nodes,
                   int32 *succ, int cap_links, int32 *segs, int cap_segs, int32 *info, int32 *lm_tables)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    allphone_search_t *ap;
    bin_mdef_t *mdef;
    phmm_t **nodes, *p;
    gnode_t *gn;
    int n_nodes = 0, n_links = 0, ci, i, n;

    err_set_loglevel(ERR_ERROR);
    config = ps_config_init(NULL);
    ps_config_set_str(config, "hmm", hmmdir);
    ps_config_set_str(config, "dither", "no");
    ps_config_set_str(config, "compallsen", "yes");
    ps_config_set_str(config, "pl_window", "0");
    ps_config_set_str(config, "lm", NULL);
    ps_config_set_str(config, "dict", NULL);
    if (kv) {
        char *b2 = strdup(kv), *s2 = NULL, *t2;
        for (t2 = strtok_r(b2, "\n", &s2); t2; t2 = strtok_r(NULL, "\n", &s2)) {
            char *eq = strchr(t2, '=');
            if (!eq) continue;
            *eq = 0;
            ps_config_set_str(config, t2, eq + 1);
        }
        free(b2);
    }
    ps = ps_init(config);
    if (ps == NULL) { ps_config_free(config); return -1; }
    if (ps->search == NULL || strcmp(ps_search_type(ps->search), PS_SEARCH_TYPE_ALLPHONE) != 0) {
        /* no "allphone" LM in the configuration: unconstrained loop */
        if (ps_add_allphone(ps, "_ap", NULL) < 0 || ps_activate_search(ps, "_ap") < 0) {
            ps_free(ps); ps_config_free(config);
            return -2;
        }
    }
    ap = (allphone_search_t *)ps->search;
    mdef = ps->acmod->mdef;
    ps_start_utt(ps);
    ps_process_raw(ps, pcm, n_samples, FALSE, TRUE);
    ps_end_utt(ps);
    for (ci = 0; ci < bin_mdef_n_ciphone(mdef); ++ci)
        for (p = ap->ci_phmm[ci]; p; p = p->next) ++n_nodes;
    nodes = calloc(n_nodes > 0 ? n_nodes : 1, sizeof(*nodes));
    for (ci = 0, i = 0; ci < bin_mdef_n_ciphone(mdef); ++ci)
        for (p = ap->ci_phmm[ci]; p; p = p->next) nodes[i++] = p;
    info[3] = -1;
    succ_off[0] = 0;
    for (i = 0; i < n_nodes; ++i) {
        plink_t *l;
        if (i < cap_nodes) {
            node_ci[i] = nodes[i]->ci; node_ssid[i] = hmm_nonmpx_ssid(&nodes[i]->hmm); node_tmat[i] = nodes[i]->hmm.tmatid;
        }
        if (nodes[i]->ci == bin_mdef_silphone(mdef) && nodes[i]->pid == bin_mdef_silphone(mdef) && info[3] < 0) info[3] = i;
        for (l = nodes[i]->succlist; l; l = l->next) {
            int j;
            for (j = 0; j < n_nodes && nodes[j] != l->phmm; ++j) ;
            if (n_links < cap_links) succ[n_links] = j;
            ++n_links;
        }
        if (i + 1 <= cap_nodes) succ_off[i + 1] = n_links;
    }
    info[0] = ps_get_n_frames(ps); info[1] = n_nodes; info[2] = n_links;
    info[4] = ap->beam; info[5] = ap->pbeam; info[6] = ap->inspen;
    n = 0;
    for (gn = ap->segments; gn; gn = gnode_next(gn)) {
        phseg_t *s = gnode_ptr(gn);
        if (n < cap_segs) {
            segs[n * 5 + 0] = s->ci; segs[n * 5 + 1] = s->sf; segs[n * 5 + 2] = s->ef;
            segs[n * 5 + 3] = s->score; segs[n * 5 + 4] = s->tscore;
        }
        ++n;
    }
    info[7] = n;
    info[8] = (int32)blkarray_list_n_valid(ap->history);
    info[9] = ap->lm != NULL;
    info[10] = bin_mdef_n_ciphone(mdef);
    if (ap->lm && lm_tables) {
        const int nc = bin_mdef_n_ciphone(mdef);
        int a, b, c3;
        int32 n_used;
        for (a = 0; a < nc; ++a)
            for (b = 0; b < nc; ++b) {
                lm_tables[a * nc + b] = ngram_bg_score(ap->lm, ap->ci2lmwid[a], ap->ci2lmwid[b], &n_used) >> SENSCR_SHIFT;
                for (c3 = 0; c3 < nc; ++c3)
                    lm_tables[nc * nc + (a * nc + b) * nc + c3] =
                        ngram_tg_score(ap->lm, ap->ci2lmwid[a], ap->ci2lmwid[b], ap->ci2lmwid[c3], &n_used) >> SENSCR_SHIFT;
            }
    }
    free(nodes);
    ps_free(ps);
    ps_config_free(config);
    return 0;
}

/* Grammar decoding through the reference's own fsg_search (fsg_search.c) on one utterance, with
 * everything the search works on flattened for the oracle:
 *   i32 block layout (all int32, returned through `blob`, sizes in info):
 *     pnodes  [n_pnode][16]: ssid, tmatid, next (succ pnode id, or link id for leaves, -1), sibling id,
 *                            logs2prob, ci_ext, ppos, leaf, ctxt.bv[8]
 *     roots   [n_state]    : first root pnode of the state (-1 none)
 *     links   [n_link][5]  : from_state, to_state, wid, logs2prob, all_ctxt (filler or single-phone word)
 *     nulloff [n_state+1], nullarc [n_null]: null arcs leaving each state, as link ids, in the
 *                            order fsg_model_arcs iterates them
 *     hist    [n_hist][13] : link id (-1 = the dummy start entry), frame, score, pred, lc, rc.bv[8]
 *   info: [0] frames [1] n_pnode [2] n_state [3] n_link [4] n_null [5] n_hist [6] beam [7] pbeam
 *         [8] wbeam [9] maxhmmpf [10] silcipid [11] n_ciphone [12] start_state [13] final_state
 *         [14] hyp score
 * vocab: the grammar's word strings by wid, newline separated.
 * Pnode ids follow alloc_head[s] / alloc_next, state by state. */
#include "fsg_search_internal.h"
#include "fsg_lextree.h"
#include "fsg_history.h"
#include "ps_search_cuda.h"

long
refdrv_fsg(const char *hmmdir, const char *dict, const char *fsgfile, const char *kv,
           const int16 *pcm, long n_samples, int32 *blob, long cap, int32 *info, char *hyp, int hyp_cap,
           char *vocab, int vocab_cap)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    fsg_search_t *fs;
    fsg_model_t *fsg;
    cuda_fsg_graph_t g;
    int n_pn, n_link, n_state, i, n_null, n_hist;
    long need, o;
    const char *h;
    int32 score = 0;

    err_set_loglevel(ERR_ERROR);
    config = ps_config_init(NULL);
    ps_config_set_str(config, "hmm", hmmdir);
    ps_config_set_str(config, "dict", dict);
    ps_config_set_str(config, "fsg", fsgfile);
    ps_config_set_str(config, "lm", NULL);
    ps_config_set_str(config, "dither", "no");
    ps_config_set_str(config, "compallsen", "yes");
    ps_config_set_str(config, "pl_window", "0");
    ps_config_set_str(config, "bestpath", "no");
    if (kv) {
        char *b2 = strdup(kv), *s2 = NULL, *t2;
        for (t2 = strtok_r(b2, "\n", &s2); t2; t2 = strtok_r(NULL, "\n", &s2)) {
            char *eq = strchr(t2, '=');
            if (!eq) continue;
            *eq = 0;
            ps_config_set_str(config, t2, eq + 1);
        }
        free(b2);
    }
    ps = ps_init(config);
    if (ps == NULL) { ps_config_free(config); return -1; }
    if (ps->search == NULL || strcmp(ps_search_type(ps->search), PS_SEARCH_TYPE_FSG) != 0) {
        ps_free(ps); ps_config_free(config);
        return -2;
    }
    fs = (fsg_search_t *)ps->search;
    ps_start_utt(ps);
    ps_process_raw(ps, pcm, n_samples, FALSE, TRUE);
    ps_end_utt(ps);
    h = ps_get_hyp(ps, &score);
    {   /* first line: the hypothesis; then one "word sf ef ascr lscr" line per ps_seg_iter segment */
        int len = snprintf(hyp, hyp_cap, "%s\n", h ? h : "");
        ps_seg_t *it;
        for (it = ps_seg_iter(ps); it && len < hyp_cap; it = ps_seg_next(it)) {
            int sf, ef;
            int32 ascr, lscr, lback;
            ps_seg_frames(it, &sf, &ef);
            ps_seg_prob(it, &ascr, &lscr, &lback);
            len += snprintf(hyp + len, hyp_cap - len, "%s %d %d %d %d\n", ps_seg_word(it), sf, ef, ascr, lscr);
        }
        if (it) ps_seg_free(it);
    }
    fsg = fs->fsg;
    if (vocab && vocab_cap > 0) {                     /* word strings by wid, newline separated */
        int w, len = 0;
        vocab[0] = 0;
        for (w = 0; w < fsg_model_n_word(fsg); ++w)
            len += snprintf(vocab + len, len < vocab_cap ? vocab_cap - len : 0, "%s\n", fsg_model_word_str(fsg, w));
    }
    /* the flattening itself is the maintainer-side binding: integration/ps_search_cuda.c */
    if (cuda_fsg_export(fs, &g) != 0) { ps_free(ps); ps_config_free(config); return -3; }
    n_pn = g.desc.n_pnode; n_state = g.desc.n_state; n_link = g.desc.n_link; n_null = g.nulloff[n_state];
    n_hist = fsg_history_n_entries(fs->history);
    need = (long)n_pn * 16 + n_state + (long)n_link * 5 + (n_state + 1) + n_null + (long)n_hist * 13;
    info[0] = ps_get_n_frames(ps); info[1] = n_pn; info[2] = n_state; info[3] = n_link; info[4] = n_null;
    info[5] = n_hist; info[6] = g.desc.beam; info[7] = g.desc.pbeam; info[8] = g.desc.wbeam;
    info[9] = g.desc.maxhmmpf; info[10] = g.desc.silcipid; info[11] = g.desc.n_ciphone;
    info[12] = g.desc.start_state; info[13] = fsg_model_final_state(fsg); info[14] = score;
    info[15] = -1;
    if (blob && cap >= need) {
        o = 0;
        memcpy(blob + o, g.pnodes, (size_t)n_pn * 16 * sizeof(int32)); o += (long)n_pn * 16;
        memcpy(blob + o, g.roots, (size_t)n_state * sizeof(int32)); o += n_state;
        memcpy(blob + o, g.links, (size_t)n_link * 5 * sizeof(int32)); o += (long)n_link * 5;
        memcpy(blob + o, g.nulloff, ((size_t)n_state + 1) * sizeof(int32)); o += n_state + 1;
        memcpy(blob + o, g.nullarc, (size_t)n_null * sizeof(int32)); o += n_null;
        for (i = 0; i < n_hist; ++i) {
            fsg_hist_entry_t *e = fsg_history_entry_get(fs->history, i);
            int32 *r = blob + o + (long)i * 13;
            int j, id = -1;
            if (e->fsglink) for (j = 0; j < g.n_link; ++j) if (g.link_ptr[j] == e->fsglink) { id = j; break; }
            if (e->fsglink && id < 0) need = -4;              /* a history link the lextree does not know */
            r[0] = id;
            r[1] = e->frame; r[2] = e->score; r[3] = e->pred; r[4] = e->lc;
            for (j = 0; j < 8; ++j) r[5 + j] = (int32)e->rc.bv[j];
        }
    }
    cuda_fsg_free(&g);
    ps_free(ps);
    ps_config_free(config);
    return need;
}

/* N-gram lextree decoding, first pass only: the reference's own ngram_search_fwdtree (fwdflat and
 * bestpath off, no phone-loop look-ahead, all senones) on one utterance, with everything that pass
 * works on flattened for the oracle.  int32 sections, in this order (sizes from info):
 *   roots    [n_root][5]      ciphone, ci2phone, penult_phn_wid, next (non-root id or -1), tmatid
 *   nonroot  [n_nonroot][6]   ssid, tmatid, ciphone, penult_phn_wid, next, alt   (ids: depth-first from the roots)
 *   words    [n_words][8]     first phone, last phone, second-last phone (-1: single phone), single-phone,
 *                             filler, basewid, homophone_set[w], index into the LM table (-1: not a base word)
 *   w1ph     [n_1ph_words]    single_phone_wid[]
 *   r1ph     [n_1ph_words][4] their permanent root channels: ciphone, ci2phone, ssid, tmatid
 *   rs_n     [n_ci][n_ci]     dict2pid rssid(last, second-last): n_ssid
 *   rs_ssid  [n_ci][n_ci][n_ci]   ssid[] padded with -1;   rs_cimap [n_ci][n_ci][n_ci]
 *   ldiph    [n_ci][n_ci][n_ci]   ldiph_lc[b][r][l]
 *   lm       [n_lm][n_lm+1][n_lm+1]  ngram_tg_score(w, h1, h2) >> SENSCR_SHIFT, history index 0 = none (-1)
 *   inlm     [n_words]        ngram_model_set_known_wid(basewid)
 *   pron_off [n_words+1], pron_ci [n_pron], pron_ssid [n_pron]: pronunciations; dict2pid_internal ssid for
 *                             word-internal positions 1..len-2, -1 elsewhere
 *   bp       [bpidx][10]      frame, valid, wid, bp, score, s_idx, real_wid, prev_real_wid, last_phone, last2_phone
 *   bss      [bss_head]       bscore_stack
 *   bpidx_f  [n_frame+1]      bp_table_idx
 * info: 0 n_frame 1 n_words 2 n_root 3 n_nonroot 4 n_1ph_words 5 n_1ph_LMwords 6 n_ci 7 silence phone
 *   8 beam 9 pbeam 10 wbeam 11 lpbeam 12 lponlybeam 13 maxhmmpf 14 maxwpf 15 nwpen 16 pip 17 silpen
 *   18 fillpen 19 start wid 20 finish wid 21 silence wid 22 filler_start 23 filler_end 24 bpidx
 *   25 bss_head 26 n_lm 27 hyp score 28 fwdflatbeam 29 fwdflatwbeam 30 fwdflatefwid 31 fwdflatsfwin
 *   32 fwdflat_fwdtree_lw_ratio (float32 bits) 33 n_pron   (info holds 40 ints)
 * With fwdflat=yes in kv the tables returned are those of the second pass (ngram_search_fwdflat.c). */
#include "ngram_search.h"
#include "ngram_search_fwdtree.h"
static ps_decoder_t *
ngram_decoder(const char *hmmdir, const char *lm, const char *dictfile, const char *kv, ps_config_t **cfg)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    err_set_loglevel(ERR_ERROR);
    config = ps_config_init(NULL);
    ps_config_set_str(config, "hmm", hmmdir);
    ps_config_set_str(config, "lm", lm);
    ps_config_set_str(config, "dict", dictfile);
    ps_config_set_str(config, "dither", "no");
    ps_config_set_str(config, "compallsen", "yes");
    ps_config_set_str(config, "pl_window", "0");
    ps_config_set_str(config, "fwdflat", "no");
    ps_config_set_str(config, "bestpath", "no");
    if (kv) {
        char *b2 = strdup(kv), *s2 = NULL, *t2;
        for (t2 = strtok_r(b2, "\n", &s2); t2; t2 = strtok_r(NULL, "\n", &s2)) {
            char *eq = strchr(t2, '=');
            if (!eq) continue;
            *eq = 0;
            ps_config_set_str(config, t2, eq + 1);
        }
        free(b2);
    }
    ps = ps_init(config);
    if (ps == NULL) { ps_config_free(config); return NULL; }
    if (ps->search == NULL || strcmp(ps_search_type(ps->search), PS_SEARCH_TYPE_NGRAM) != 0) {
        ps_free(ps); ps_config_free(config);
        return NULL;
    }
    *cfg = config;
    return ps;
}

long
refdrv_fwdtree(const char *hmmdir, const char *lm, const char *dictfile, const char *kv,
               const int16 *pcm, long n_samples, int32 *blob, long cap, int32 *info, char *hyp, int hyp_cap,
               char *vocab, int vocab_cap, int dense_lm)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    ngram_search_t *ngs;
    dict_t *dict;
    cuda_ngram_graph_t g;
    int i, w, n_words;
    long need, o;
    const char *h;
    int32 score = 0;

    if ((ps = ngram_decoder(hmmdir, lm, dictfile, kv, &config)) == NULL) return -1;
    ngs = (ngram_search_t *)ps->search;
    ps_start_utt(ps);
    ps_process_raw(ps, pcm, n_samples, FALSE, TRUE);
    ps_end_utt(ps);
    h = ps_get_hyp(ps, &score);
    snprintf(hyp, hyp_cap, "%s", h ? h : "");
    dict = ps_search_dict(ngs);
    n_words = ps_search_n_words(ngs);
    if (vocab && vocab_cap > 0) {                     /* word strings by wid, newline separated */
        int len = 0;
        vocab[0] = 0;
        for (w = 0; w < n_words; ++w)
            len += snprintf(vocab + len, len < vocab_cap ? vocab_cap - len : 0, "%s\n", dict_wordstr(dict, w));
    }
    /* the flattening itself is the maintainer-side binding: integration/ps_search_cuda.c */
    if (cuda_ngram_export(ngs, &g, dense_lm) != 0) { ps_free(ps); ps_config_free(config); return -3; }
    need = (long)g.model_len + (long)ngs->bpidx * 10 + ngs->bss_head + ngs->n_frame + 1;
    memcpy(info, g.info, 40 * sizeof(int32));
    info[0] = ngs->n_frame; info[24] = ngs->bpidx; info[25] = ngs->bss_head; info[27] = score;
    if (blob && cap >= need) {
        memcpy(blob, g.model, (size_t)g.model_len * sizeof(int32));
        o = (long)g.model_len;
        for (i = 0; i < ngs->bpidx; ++i) {
            bptbl_t *b = &ngs->bp_table[i];
            blob[o++] = b->frame; blob[o++] = b->valid; blob[o++] = b->wid; blob[o++] = b->bp; blob[o++] = b->score;
            blob[o++] = b->s_idx; blob[o++] = b->real_wid; blob[o++] = b->prev_real_wid; blob[o++] = b->last_phone;
            blob[o++] = b->last2_phone;
        }
        for (i = 0; i < ngs->bss_head; ++i) blob[o++] = ngs->bscore_stack[i];
        for (i = 0; i <= ngs->n_frame; ++i) blob[o++] = ngs->bp_table_idx[i];
        if (o != need) need = -3;
    }
    cuda_ngram_free(&g);
    ps_free(ps);
    ps_config_free(config);
    return need;
}

/* Round trips through the binding: decode normally, WIPE the search's own result tables, import tables
 * computed elsewhere (the oracle, the host-emulated phase code, the device) with cuda_*_import, and let
 * the reference's unchanged code (fsg_search_hyp; ngram_search_hyp with lattice + bestpath when
 * configured) produce hypothesis and score from them. */
long
refdrv_fsg_roundtrip(const char *hmmdir, const char *dict, const char *fsgfile, const char *kv,
                     const int16 *pcm, long n_samples, const int32 *rows, int32 n_rows, int32 n_frames,
                     char *hyp, int hyp_cap, int32 *score_out)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    fsg_search_t *fs;
    cuda_fsg_graph_t g;
    const char *h;
    int32 score = 0;
    long rv = 0;

    err_set_loglevel(ERR_ERROR);
    config = ps_config_init(NULL);
    ps_config_set_str(config, "hmm", hmmdir);
    ps_config_set_str(config, "dict", dict);
    ps_config_set_str(config, "fsg", fsgfile);
    ps_config_set_str(config, "lm", NULL);
    ps_config_set_str(config, "dither", "no");
    ps_config_set_str(config, "compallsen", "yes");
    ps_config_set_str(config, "pl_window", "0");
    ps_config_set_str(config, "bestpath", "no");
    if (kv) {
        char *b2 = strdup(kv), *s2 = NULL, *t2;
        for (t2 = strtok_r(b2, "\n", &s2); t2; t2 = strtok_r(NULL, "\n", &s2)) {
            char *eq = strchr(t2, '=');
            if (!eq) continue;
            *eq = 0;
            ps_config_set_str(config, t2, eq + 1);
        }
        free(b2);
    }
    ps = ps_init(config);
    if (ps == NULL) { ps_config_free(config); return -1; }
    if (ps->search == NULL || strcmp(ps_search_type(ps->search), PS_SEARCH_TYPE_FSG) != 0) { ps_free(ps); ps_config_free(config); return -2; }
    fs = (fsg_search_t *)ps->search;
    ps_start_utt(ps);
    ps_process_raw(ps, pcm, n_samples, FALSE, TRUE);
    ps_end_utt(ps);
    if (cuda_fsg_export(fs, &g) != 0) { ps_free(ps); ps_config_free(config); return -3; }
    if (cuda_fsg_import(fs, &g, rows, n_rows, n_frames) != 0) rv = -4;
    else {
        h = ps_get_hyp(ps, &score);
        snprintf(hyp, hyp_cap, "%s", h ? h : "");
        *score_out = score;
        rv = fsg_history_n_entries(fs->history);
    }
    cuda_fsg_free(&g);
    ps_free(ps);
    ps_config_free(config);
    return rv;
}

long
refdrv_ngram_roundtrip(const char *hmmdir, const char *lm, const char *dictfile, const char *kv,
                       const int16 *pcm, long n_samples, const int32 *bp, int32 n_bp, const int32 *bss, int32 n_bss,
                       const int32 *bp_idx, int32 n_frames, char *hyp, int hyp_cap, int32 *score_out,
                       char *seg, int seg_cap)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    ngram_search_t *ngs;
    const char *h;
    int32 score = 0;
    long rv;

    if ((ps = ngram_decoder(hmmdir, lm, dictfile, kv, &config)) == NULL) return -1;
    ngs = (ngram_search_t *)ps->search;
    ps_start_utt(ps);
    ps_process_raw(ps, pcm, n_samples, FALSE, TRUE);
    ps_end_utt(ps);
    /* wipe, so that nothing of the reference's own pass can leak into the result */
    memset(ngs->bp_table, 0xff, (size_t)ngs->bp_table_size * sizeof(*ngs->bp_table));
    memset(ngs->bscore_stack, 0xff, (size_t)ngs->bscore_stack_size * sizeof(*ngs->bscore_stack));
    if (cuda_ngram_import(ngs, bp, n_bp, bss, n_bss, bp_idx, n_frames) != 0) rv = -4;
    else {
        ps_seg_t *it;
        int len = 0;
        h = ps_get_hyp(ps, &score);
        snprintf(hyp, hyp_cap, "%s", h ? h : "");
        *score_out = score;
        if (seg && seg_cap > 0) {
            seg[0] = 0;
            for (it = ps_seg_iter(ps); it; it = ps_seg_next(it)) {
                int sf, ef;
                ps_seg_frames(it, &sf, &ef);
                len += snprintf(seg + len, len < seg_cap ? seg_cap - len : 0, "%s %d %d\n", ps_seg_word(it), sf, ef);
            }
        }
        rv = ngs->bpidx;
    }
    ps_free(ps);
    ps_config_free(config);
    return rv;
}

/* The language model behind an n-gram search as sorted arrays (cuda_ngram_export_lm), plus a sample of
 * the reference's own scores for checking: for every (w, h1, h2) in `q` [n_q][3] (dictionary word ids, -1 =
 * no history) ngram_tg_score(lmset, w, h1, h2) >> SENSCR_SHIFT is written to scores[n_q]. */
long
refdrv_lm_arrays(const char *hmmdir, const char *lm, const char *dictfile, const char *kv, int32 *out, long cap,
                 const int32 *q, long n_q, int32 *scores)
{
    ps_config_t *config;
    ps_decoder_t *ps;
    ngram_search_t *ngs;
    long need, i;
    if ((ps = ngram_decoder(hmmdir, lm, dictfile, kv, &config)) == NULL) return -1;
    ngs = (ngram_search_t *)ps->search;
    need = cuda_ngram_export_lm(ngs, out, cap);
    for (i = 0; i < n_q && scores; ++i) {
        int32 n_used;
        scores[i] = ngram_tg_score(ngs->lmset, q[i * 3], q[i * 3 + 1], q[i * 3 + 2], &n_used) >> SENSCR_SHIFT;
    }
    ps_free(ps);
    ps_config_free(config);
    return need;
}
