/* integration/ps_search_cuda.c -- reference-side binding for the device search modules of
 * libpsb200.so (psb_fsg_batch_device, psb_ngram_fwdtree_batch_device, psb_ngram_fwdflat_batch_device).
 *
 * Like integration/ps_mgau_cuda.c this is a file a PocketSphinx maintainer adds to src/: it compiles
 * against the reference's internal headers.  The search objects keep doing what they do on the host
 * -- fsg_search_init / fsg_lextree_init, ngram_search_init / ngram_fwdtree_init build the lextrees,
 * read the dictionary, dict2pid and the language model -- and this file
 *
 *   1. flattens what they built into the plain int32 arrays include/psb200.h describes
 *      (cuda_fsg_export, cuda_ngram_export), once per grammar / language model, and
 *   2. after the device has searched a batch, puts an utterance's tables back where the reference
 *      keeps them (cuda_fsg_import: fsg_history_t; cuda_ngram_import: bp_table / bscore_stack /
 *      bp_table_idx), so that fsg_search_hyp, fsg_search_seg_iter, fsg_search_lattice,
 *      ngram_search_hyp, ngram_search_lattice and ps_lattice_bestpath run unchanged on them.
 *
 * In this repository the file is compiled into the test copy of the reference
 * (oracle/_ref/libpsref.so); oracle/ref_driver.c goes through it for every grammar / n-gram test and
 * tests/test_search_binding.py decodes, wipes the reference's tables, imports tables computed
 * outside the reference and requires the reference's own hypothesis, score and (with bestpath) the
 * lattice result to come out the same.
 */
#include <stdlib.h>
#include <string.h>

#include <pocketsphinx.h>

#include "fsg_search_internal.h"
#include "fsg_lextree.h"
#include "fsg_history.h"
#include "ngram_search.h"
#include "ngram_search_fwdtree.h"
#include "dict2pid.h"
#include "util/ckd_alloc.h"
#include "util/blkarray_list.h"

#include "psb200.h"
#include "ps_search_cuda.h"

/* ------------------------------------------------------------------------------------------ */
/* grammar search                                                                               */

static int
link_id(cuda_fsg_graph_t *g, fsg_link_t *l)
{
    int i;
    for (i = 0; i < g->n_link; ++i)
        if (g->link_ptr[i] == l) return i;
    g->link_ptr[g->n_link] = l;
    return g->n_link++;
}

int
cuda_fsg_export(fsg_search_t *fs, cuda_fsg_graph_t *g)
{
    fsg_lextree_t *lt = fs->lextree;
    fsg_model_t *fsg = fs->fsg;
    dict_t *dict = ps_search_dict(fs);
    fsg_pnode_t **pn, *p;
    int n_pn = 0, n_state = fsg_model_n_state(fsg), s, i, j, k, n_null = 0, n_arc = 0;

    memset(g, 0, sizeof(*g));
    for (s = 0; s < n_state; ++s)
        for (p = lt->alloc_head[s]; p; p = p->alloc_next) ++n_pn;
    pn = ckd_calloc(n_pn > 0 ? n_pn : 1, sizeof(*pn));
    for (s = 0, i = 0; s < n_state; ++s)
        for (p = lt->alloc_head[s]; p; p = p->alloc_next) pn[i++] = p;
    for (s = 0; s < n_state; ++s) {
        fsg_arciter_t *it;
        for (it = fsg_model_arcs(fsg, s); it; it = fsg_arciter_next(it)) {
            ++n_arc;
            if (fsg_link_wid(fsg_arciter_get(it)) == -1) ++n_null;
        }
    }
    g->link_ptr = ckd_calloc((size_t)n_arc + 1, sizeof(*g->link_ptr));
    /* null arcs first (their ids are what null_prop walks), then the word links leaves end in */
    g->nulloff = ckd_calloc(n_state + 1, sizeof(int32));
    g->nullarc = ckd_calloc(n_null > 0 ? n_null : 1, sizeof(int32));
    for (s = 0, k = 0; s < n_state; ++s) {
        fsg_arciter_t *it;
        g->nulloff[s] = k;
        for (it = fsg_model_arcs(fsg, s); it; it = fsg_arciter_next(it))
            if (fsg_link_wid(fsg_arciter_get(it)) == -1) g->nullarc[k++] = link_id(g, fsg_arciter_get(it));
    }
    g->nulloff[n_state] = k;
    g->pnodes = ckd_calloc((size_t)(n_pn > 0 ? n_pn : 1) * 16, sizeof(int32));
    for (i = 0; i < n_pn; ++i) {
        int32 *r = g->pnodes + (size_t)i * 16;
        p = pn[i];
        r[0] = hmm_nonmpx_ssid(&p->hmm); r[1] = p->hmm.tmatid;
        if (p->leaf) r[2] = link_id(g, p->next.fsglink);
        else {
            r[2] = -1;
            for (j = 0; j < n_pn; ++j) if (pn[j] == p->next.succ) { r[2] = j; break; }
        }
        r[3] = -1;
        for (j = 0; j < n_pn; ++j) if (pn[j] == p->sibling) { r[3] = j; break; }
        r[4] = p->logs2prob; r[5] = p->ci_ext; r[6] = p->ppos; r[7] = p->leaf;
        for (j = 0; j < FSG_PNODE_CTXT_BVSZ && j < 8; ++j) r[8 + j] = (int32)p->ctxt.bv[j];
    }
    g->roots = ckd_calloc(n_state, sizeof(int32));
    for (s = 0; s < n_state; ++s) {
        g->roots[s] = -1;
        for (i = 0; i < n_pn; ++i) if (pn[i] == lt->root[s]) { g->roots[s] = i; break; }
    }
    g->links = ckd_calloc((size_t)(g->n_link > 0 ? g->n_link : 1) * 5, sizeof(int32));
    for (i = 0; i < g->n_link; ++i) {
        fsg_link_t *l = g->link_ptr[i];
        int32 *r = g->links + (size_t)i * 5;
        r[0] = l->from_state; r[1] = l->to_state; r[2] = l->wid; r[3] = l->logs2prob;
        r[4] = 0;
        if (l->wid >= 0)                                      /* fsg_search_pnode_exit, fsg_search.c:468-474 */
            r[4] = fsg_model_is_filler(fsg, l->wid)
                || dict_is_single_phone(dict, dict_wordid(dict, fsg_model_word_str(fsg, l->wid)));
    }
    ckd_free(pn);
    g->desc.n_pnode = n_pn; g->desc.pnodes = g->pnodes;
    g->desc.n_state = n_state; g->desc.roots = g->roots;
    g->desc.n_link = g->n_link; g->desc.links = g->links;
    g->desc.nulloff = g->nulloff; g->desc.nullarc = g->nullarc;
    g->desc.n_ciphone = bin_mdef_n_ciphone(ps_search_acmod(fs)->mdef);
    g->desc.silcipid = bin_mdef_ciphone_id(ps_search_acmod(fs)->mdef, "SIL");
    g->desc.start_state = fsg_model_start_state(fsg);
    g->desc.beam = fs->beam_orig; g->desc.pbeam = fs->pbeam_orig; g->desc.wbeam = fs->wbeam_orig;
    g->desc.maxhmmpf = ps_config_int(ps_search_config(fs), "maxhmmpf");
    return 0;
}

void
cuda_fsg_free(cuda_fsg_graph_t *g)
{
    ckd_free(g->pnodes); ckd_free(g->roots); ckd_free(g->links); ckd_free(g->nulloff); ckd_free(g->nullarc);
    ckd_free(g->link_ptr);
    memset(g, 0, sizeof(*g));
}

/* rows [n][13] = link, frame, score, pred, lc, rc.bv[8] (psb_fsg_batch_device): they become the
 * utterance's fsg_history_t, as if fsg_search_step had run n_frames times and fsg_search_finish after it. */
int
cuda_fsg_import(fsg_search_t *fs, const cuda_fsg_graph_t *g, const int32 *rows, int32 n, int32 n_frames)
{
    int32 i, j;
    fsg_history_reset(fs->history);
    for (i = 0; i < n; ++i) {
        const int32 *r = rows + (size_t)i * 13;
        fsg_hist_entry_t *e;
        if (r[0] < -1 || r[0] >= g->n_link || r[3] < -1 || r[3] >= i) return -1;
        e = ckd_calloc(1, sizeof(*e));
        e->fsglink = r[0] < 0 ? NULL : g->link_ptr[r[0]];
        e->frame = r[1]; e->score = r[2]; e->pred = r[3]; e->lc = (int16)r[4];
        for (j = 0; j < FSG_PNODE_CTXT_BVSZ && j < 8; ++j) e->rc.bv[j] = (uint32)r[5 + j];
        blkarray_list_append(fs->history->entries, (void *)e);
    }
    fs->frame = n_frames;
    fs->final = TRUE;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* n-gram search                                                                                */

static int
count_chan(chan_t *h)
{
    int n = 0;
    for (; h; h = h->alt) n += 1 + count_chan(h->next);
    return n;
}
static void
collect_chan(chan_t *h, chan_t **tab, int *n)
{
    for (; h; h = h->alt) { tab[(*n)++] = h; collect_chan(h->next, tab, n); }
}
static int
chan_id(chan_t **tab, int n, chan_t *h)
{
    int i;
    if (h == NULL) return -1;
    for (i = 0; i < n; ++i) if (tab[i] == h) return i;
    return -2;
}

int
cuda_ngram_export(ngram_search_t *ngs, cuda_ngram_graph_t *g, int dense_lm)
{
    dict_t *dict = ps_search_dict(ngs);
    dict2pid_t *d2p = ps_search_dict2pid(ngs);
    bin_mdef_t *mdef = ps_search_acmod(ngs)->mdef;
    chan_t **tab;
    int32 *lmidx, *rev, *b, *info = g->info;
    int n_nonroot = 0, n_ci = bin_mdef_n_ciphone(mdef), n_words = ps_search_n_words(ngs), n_lm = 0, n_pron = 0, i, j, k, w;
    size_t need, o = 0;

    memset(g, 0, sizeof(*g));
    info = g->info;
    for (i = 0; i < ngs->n_root_chan; ++i) n_nonroot += count_chan(ngs->root_chan[i].next);
    tab = ckd_calloc(n_nonroot + 1, sizeof(*tab));
    k = 0;
    for (i = 0; i < ngs->n_root_chan; ++i) collect_chan(ngs->root_chan[i].next, tab, &k);
    lmidx = ckd_calloc(n_words, sizeof(*lmidx));
    /* dense_lm == 0: no trigram table (vocabularies beyond a few hundred words: cuda_ngram_export_lm instead) */
    for (w = 0; w < n_words; ++w) lmidx[w] = (dense_lm && dict_basewid(dict, w) == w) ? n_lm++ : -1;
    for (w = 0; w < n_words; ++w) n_pron += dict_pronlen(dict, w);
    need = (size_t)ngs->n_root_chan * 5 + (size_t)n_nonroot * 6 + (size_t)n_words * 8 + (size_t)ngs->n_1ph_words * 5
        + (size_t)n_ci * n_ci + 3 * (size_t)n_ci * n_ci * n_ci + (size_t)n_lm * (n_lm + 1) * (n_lm + 1)
        + (size_t)n_words + (n_words + 1) + 2 * (size_t)n_pron;
    b = g->model = ckd_calloc(need, sizeof(int32));
    g->model_len = (int64_t)need;
    info[1] = n_words; info[2] = ngs->n_root_chan; info[3] = n_nonroot;
    info[4] = ngs->n_1ph_words; info[5] = ngs->n_1ph_LMwords; info[6] = n_ci; info[7] = mdef->sil;
    info[8] = ngs->beam; info[9] = ngs->pbeam; info[10] = ngs->wbeam; info[11] = ngs->lpbeam; info[12] = ngs->lponlybeam;
    info[13] = ngs->maxhmmpf; info[14] = ngs->maxwpf; info[15] = ngs->nwpen; info[16] = ngs->pip;
    info[17] = ngs->silpen; info[18] = ngs->fillpen; info[19] = dict_startwid(dict); info[20] = ps_search_finish_wid(ngs);
    info[21] = ps_search_silence_wid(ngs); info[22] = dict_filler_start(dict); info[23] = dict_filler_end(dict);
    info[26] = n_lm;
    info[28] = ngs->fwdflatbeam; info[29] = ngs->fwdflatwbeam; info[30] = ngs->min_ef_width; info[31] = ngs->max_sf_win;
    memcpy(&info[32], &ngs->fwdflat_fwdtree_lw_ratio, 4); info[33] = n_pron;
    for (i = 0; i < ngs->n_root_chan; ++i) {
        root_chan_t *r = &ngs->root_chan[i];
        b[o++] = r->ciphone; b[o++] = r->ci2phone; b[o++] = r->penult_phn_wid;
        b[o++] = chan_id(tab, n_nonroot, r->next); b[o++] = r->hmm.tmatid;
    }
    for (i = 0; i < n_nonroot; ++i) {
        chan_t *c = tab[i];
        b[o++] = hmm_nonmpx_ssid(&c->hmm); b[o++] = c->hmm.tmatid; b[o++] = c->ciphone;
        b[o++] = c->info.penult_phn_wid; b[o++] = chan_id(tab, n_nonroot, c->next); b[o++] = chan_id(tab, n_nonroot, c->alt);
    }
    for (w = 0; w < n_words; ++w) {
        b[o++] = dict_first_phone(dict, w); b[o++] = dict_last_phone(dict, w);
        b[o++] = dict_is_single_phone(dict, w) ? -1 : dict_second_last_phone(dict, w);
        b[o++] = dict_is_single_phone(dict, w); b[o++] = dict_filler_word(dict, w);
        b[o++] = dict_basewid(dict, w); b[o++] = ngs->homophone_set ? ngs->homophone_set[w] : -1; b[o++] = lmidx[w];
    }
    for (i = 0; i < ngs->n_1ph_words; ++i) b[o++] = ngs->single_phone_wid[i];
    for (i = 0; i < ngs->n_1ph_words; ++i) {
        root_chan_t *r = (root_chan_t *)ngs->word_chan[ngs->single_phone_wid[i]];
        b[o++] = r->ciphone; b[o++] = r->ci2phone;
        b[o++] = bin_mdef_pid2ssid(mdef, r->ciphone); b[o++] = r->hmm.tmatid;
    }
    for (i = 0; i < n_ci; ++i) for (j = 0; j < n_ci; ++j) b[o++] = dict2pid_rssid(d2p, i, j)->n_ssid;
    for (i = 0; i < n_ci; ++i) for (j = 0; j < n_ci; ++j) {
        xwdssid_t *x = dict2pid_rssid(d2p, i, j);
        for (k = 0; k < n_ci; ++k) b[o++] = (x->ssid && k < x->n_ssid) ? x->ssid[k] : -1;
    }
    for (i = 0; i < n_ci; ++i) for (j = 0; j < n_ci; ++j) {
        xwdssid_t *x = dict2pid_rssid(d2p, i, j);
        for (k = 0; k < n_ci; ++k) b[o++] = x->cimap ? x->cimap[k] : -1;
    }
    for (i = 0; i < n_ci; ++i) for (j = 0; j < n_ci; ++j) for (k = 0; k < n_ci; ++k)
        b[o++] = (d2p->ldiph_lc[i] && d2p->ldiph_lc[i][j]) ? dict2pid_ldiph_lc(d2p, i, j, k) : -1;
    rev = ckd_calloc(n_lm + 1, sizeof(*rev));
    rev[0] = -1;
    for (w = 0; w < n_words; ++w) if (lmidx[w] >= 0) rev[lmidx[w] + 1] = w;
    for (i = 0; i < n_lm; ++i) for (j = 0; j <= n_lm; ++j) for (k = 0; k <= n_lm; ++k) {
        int32 n_used;
        b[o++] = ngram_tg_score(ngs->lmset, rev[i + 1], rev[j], rev[k], &n_used) >> SENSCR_SHIFT;
    }
    ckd_free(rev);
    for (w = 0; w < n_words; ++w) b[o++] = ngram_model_set_known_wid(ngs->lmset, dict_basewid(dict, w)) ? 1 : 0;
    for (w = 0, k = 0; w < n_words; ++w) { b[o++] = k; k += dict_pronlen(dict, w); }
    b[o++] = k;
    for (w = 0; w < n_words; ++w) for (j = 0; j < dict_pronlen(dict, w); ++j) b[o++] = dict_pron(dict, w, j);
    for (w = 0; w < n_words; ++w) for (j = 0; j < dict_pronlen(dict, w); ++j)
        b[o++] = (j >= 1 && j < dict_pronlen(dict, w) - 1) ? dict2pid_internal(d2p, w, j) : -1;
    ckd_free(tab); ckd_free(lmidx);
    if (o != need) return -1;
    g->ci_tmat = ckd_calloc(n_ci, sizeof(int32));
    g->ci_ssid = ckd_calloc(n_ci, sizeof(int32));
    for (i = 0; i < n_ci; ++i) { g->ci_tmat[i] = bin_mdef_pid2tmatid(mdef, i); g->ci_ssid[i] = bin_mdef_pid2ssid(mdef, i); }
    g->desc.info = g->info; g->desc.model = g->model; g->desc.model_len = g->model_len;
    g->desc.ci_tmat = g->ci_tmat; g->desc.ci_ssid = g->ci_ssid;
    return 0;
}

void
cuda_ngram_free(cuda_ngram_graph_t *g)
{
    ckd_free(g->model); ckd_free(g->ci_tmat); ckd_free(g->ci_ssid);
    memset(g, 0, sizeof(*g));
}

/* bp [n][10] (frame, valid, wid, bp, score, s_idx, real_wid, prev_real_wid, last_phone, last2_phone), the
 * score stack and bp_table_idx of one utterance (psb_ngram_fwdtree/fwdflat_batch_device) become the
 * search's tables, as after ngram_fwdtree_finish / ngram_fwdflat_finish; cached results are dropped. */
int
cuda_ngram_import(ngram_search_t *ngs, const int32 *bp, int32 n, const int32 *bss, int32 n_bss, const int32 *bp_idx,
                  int32 n_frames)
{
    ps_search_t *base = ps_search_base(ngs);
    int32 i;
    if (n > ngs->bp_table_size) {
        while (ngs->bp_table_size < n) ngs->bp_table_size *= 2;
        ngs->bp_table = ckd_realloc(ngs->bp_table, ngs->bp_table_size * sizeof(*ngs->bp_table));
    }
    if (n_bss > ngs->bscore_stack_size) {
        while (ngs->bscore_stack_size < n_bss) ngs->bscore_stack_size *= 2;
        ngs->bscore_stack = ckd_realloc(ngs->bscore_stack, ngs->bscore_stack_size * sizeof(*ngs->bscore_stack));
    }
    while (n_frames >= ngs->n_frame_alloc) {
        ngs->n_frame_alloc *= 2;
        ngs->bp_table_idx = ckd_realloc(ngs->bp_table_idx - 1, (ngs->n_frame_alloc + 1) * sizeof(*ngs->bp_table_idx));
        if (ngs->frm_wordlist)
            ngs->frm_wordlist = ckd_realloc(ngs->frm_wordlist, ngs->n_frame_alloc * sizeof(*ngs->frm_wordlist));
        ++ngs->bp_table_idx;
    }
    for (i = 0; i < n; ++i) {
        const int32 *r = bp + (size_t)i * 10;
        bptbl_t *e = &ngs->bp_table[i];
        if (r[3] < -1 || r[3] >= i || r[2] < 0 || r[2] >= ps_search_n_words(ngs)) return -1;
        memset(e, 0, sizeof(*e));
        e->frame = r[0]; e->valid = (uint8)r[1]; e->wid = r[2]; e->bp = r[3]; e->score = r[4]; e->s_idx = r[5];
        e->real_wid = r[6]; e->prev_real_wid = r[7]; e->last_phone = (int16)r[8]; e->last2_phone = (int16)r[9];
    }
    memcpy(ngs->bscore_stack, bss, (size_t)n_bss * sizeof(int32));
    memcpy(ngs->bp_table_idx, bp_idx, ((size_t)n_frames + 1) * sizeof(int32));
    ngs->bpidx = n; ngs->bss_head = n_bss; ngs->n_frame = n_frames;
    ngs->done = TRUE;
    if (base->dag) { ps_lattice_free(base->dag); base->dag = NULL; }
    base->last_link = NULL;
    base->post = 0;
    ckd_free(base->hyp_str);
    base->hyp_str = NULL;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* The language model as sorted arrays (for vocabularies where the dense trigram table of
 * cuda_ngram_export is out of reach).  One trie model of order <= 3 behind the search's model set.
 * The trie is keyed word -> nearest history word -> next history word (lm_trie.c:665-706): unigram w
 * owns the bigram entries (w | h1), each of which owns the trigram entries (w | h2 h1); inside a
 * range entries are sorted by word id.  Probabilities and backoffs are the floats the reference's own
 * readers return (dequantised), lw / log_wip / log_zero the trie model's (weight_score,
 * ngram_model_trie.c:709-713).  int32 layout:
 *   [0] order [1] V [2] n2 [3] n3 [4] lw (float bits) [5] log_wip [6] log_zero [7] n_words
 *   [8] max_vocab of the bigram level, [9] of the trigram level (uniform_find's upper key bound, lm_trie.c:566)
 *   widmap [n_words] | uni_prob [V] | uni_bo [V] | uni_next [V+1]
 *   | bg_word [n2] | bg_prob [n2] | bg_bo [n2] | bg_next [n2+1] | tg_word [n3] | tg_prob [n3]          */
#include "lm/ngram_model_set.h"
#include "lm/ngram_model_trie.h"
#include "lm/lm_trie.h"

static int32
f2i(float f) { int32 i; memcpy(&i, &f, 4); return i; }

long
cuda_ngram_export_lm(ngram_search_t *ngs, int32 *out, long cap)
{
    ngram_model_set_t *set = (ngram_model_set_t *)ngs->lmset;
    ngram_model_trie_t *tm;
    lm_trie_t *t;
    int order, V, n_words = ps_search_n_words(ngs), w;
    uint32 n2 = 0, n3 = 0, p;
    long need, o;

    if (set == NULL || set->n_models != 1 || set->cur != 0) return -1;       /* interpolated sets: not handled */
    tm = (ngram_model_trie_t *)set->lms[0];
    t = tm->trie;
    order = tm->base.n;
    V = tm->base.n_counts[0];
    if (order < 1 || order > 3 || t == NULL) return -2;
    if (order >= 2) n2 = tm->base.n_counts[1];
    if (order >= 3) n3 = tm->base.n_counts[2];
    need = 10 + n_words + 2L * V + (V + 1) + 3L * n2 + (n2 + 1) + 2L * n3;
    if (out == NULL || cap < need) return need;
    o = 0;
    out[o++] = order; out[o++] = V; out[o++] = (int32)n2; out[o++] = (int32)n3;
    out[o++] = f2i(tm->base.lw); out[o++] = tm->base.log_wip; out[o++] = tm->base.log_zero; out[o++] = n_words;
    out[o++] = order == 3 ? (int32)t->middle_begin[0].base.max_vocab : (order == 2 ? (int32)t->longest->base.max_vocab : 0);
    out[o++] = order == 3 ? (int32)t->longest->base.max_vocab : 0;
    for (w = 0; w < n_words; ++w) out[o++] = set->widmap[w][0];
    for (w = 0; w < V; ++w) out[o++] = f2i(t->unigrams[w].prob);
    for (w = 0; w < V; ++w) out[o++] = f2i(t->unigrams[w].bo);
    for (w = 0; w <= V; ++w) out[o++] = (int32)t->unigrams[w].next;
    if (order == 2) {                                      /* bigrams are the longest order: no backoff, no children */
        longest_t *l = t->longest;
        bitarr_address_t a;
        a.base = l->base.base;
        for (p = 0; p < n2; ++p) { a.offset = p * l->base.total_bits; out[o + p] = (int32)bitarr_read_int25(a, l->base.word_bits, l->base.word_mask); }
        o += n2;
        for (p = 0; p < n2; ++p) { a.offset = p * l->base.total_bits + l->base.word_bits; out[o + p] = f2i(lm_trie_quant_lpread(t->quant, a)); }
        o += n2;
        for (p = 0; p < n2; ++p) out[o + p] = 0;
        o += n2;
        for (p = 0; p <= n2; ++p) out[o + p] = 0;
        o += n2 + 1;
    }
    else if (order == 3) {
        middle_t *m = &t->middle_begin[0];
        longest_t *l = t->longest;
        bitarr_address_t a;
        a.base = m->base.base;
        for (p = 0; p < n2; ++p) { a.offset = p * m->base.total_bits; out[o + p] = (int32)bitarr_read_int25(a, m->base.word_bits, m->base.word_mask); }
        o += n2;
        for (p = 0; p < n2; ++p) { a.offset = p * m->base.total_bits + m->base.word_bits; out[o + p] = f2i(lm_trie_quant_mpread(t->quant, a, 0)); }
        o += n2;
        for (p = 0; p < n2; ++p) { a.offset = p * m->base.total_bits + m->base.word_bits; out[o + p] = f2i(lm_trie_quant_mboread(t->quant, a, 0)); }
        o += n2;
        for (p = 0; p <= n2; ++p) {
            a.offset = p * m->base.total_bits + m->base.word_bits + m->quant_bits;
            out[o + p] = (int32)bitarr_read_int25(a, m->next_mask.bits, m->next_mask.mask);
        }
        o += n2 + 1;
        a.base = l->base.base;
        for (p = 0; p < n3; ++p) { a.offset = p * l->base.total_bits; out[o + p] = (int32)bitarr_read_int25(a, l->base.word_bits, l->base.word_mask); }
        o += n3;
        for (p = 0; p < n3; ++p) { a.offset = p * l->base.total_bits + l->base.word_bits; out[o + p] = f2i(lm_trie_quant_lpread(t->quant, a)); }
        o += n3;
    }
    else { o += 3L * n2 + (n2 + 1); }
    return o == need ? need : -3;
}
