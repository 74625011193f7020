// psb_lm_core.h -- trigram scores from a language model laid out as sorted arrays: what
// ngram_tg_score(lmset, w, h1, h2) computes for one trie model of order <= 3 behind the search's
// model set (ngram_model_set_score -> ngram_ng_score -> ngram_model_trie_score -> lm_trie_score,
// lm/ngram_model_set.c:685, lm/ngram_model.c:388, lm/ngram_model_trie.c:709-742, lm/lm_trie.c:653-825),
// restated over plain arrays so that it can run on the device (and in the host harnesses).  The trie
// is keyed word -> nearest history word -> next history word; inside a node's range entries are
// meant to be sorted by word id and are searched with the reference's own interpolation search (lm_find).  Floats are
// the reference's own dequantised values; the float sum is truncated to int32 and then weighted:
// (int32)(score * lw + log_wip).  Layout of the int32 block: integration/ps_search_cuda.c
// (cuda_ngram_export_lm).  Checked against the reference on every (w, h1, h2) of the turtle and
// tidigits LMs and on 800 k queries of the 72 k-word en-us LM (tests/test_lm_arrays.py).
#pragma once
#include "psb_fsg_core.h"

struct LmArr {
    int order, V, n2, n3, log_wip, log_zero, n_words;
    uint32_t max_vocab2, max_vocab3;
    float lw;
    const int32_t *widmap, *uni_next, *bg_word, *bg_next, *tg_word;
    const float *uni_prob, *uni_bo, *bg_prob, *bg_bo, *tg_prob;
};

FSG_HDH size_t lm_arr_words(const int32_t *a)
{
    return 10 + (size_t)a[7] + 2 * (size_t)a[1] + ((size_t)a[1] + 1) + 3 * (size_t)a[2] + ((size_t)a[2] + 1) + 2 * (size_t)a[3];
}

// header values from `hdr` (host memory), array pointers relative to `base` (host or device copy of the block)
FSG_HDH void lm_arr_bind(LmArr &L, const int32_t *hdr, const int32_t *base)
{
    L.order = hdr[0]; L.V = hdr[1]; L.n2 = hdr[2]; L.n3 = hdr[3]; L.log_wip = hdr[5]; L.log_zero = hdr[6]; L.n_words = hdr[7];
    L.max_vocab2 = (uint32_t)hdr[8]; L.max_vocab3 = (uint32_t)hdr[9];
    union { int32_t i; float f; } u;
    u.i = hdr[4]; L.lw = u.f;
    const int32_t *a = base + 10;
    L.widmap = a; a += L.n_words;
    L.uni_prob = (const float *)a; a += L.V;  L.uni_bo = (const float *)a; a += L.V;  L.uni_next = a; a += L.V + 1;
    L.bg_word = a; a += L.n2;  L.bg_prob = (const float *)a; a += L.n2;  L.bg_bo = (const float *)a; a += L.n2;
    L.bg_next = a; a += L.n2 + 1;  L.tg_word = a; a += L.n3;  L.tg_prob = (const float *)a;
}

// uniform_find (lm/lm_trie.c:556-592) as it stands: interpolation search between (begin - 1, value 0) and
// (end, value max_vocab), all in uint32 arithmetic.  On a sorted range it finds exactly the entries that
// exist; shipped models (en-us.lm.bin) contain ranges that are NOT sorted, and what the reference finds
// there is a property of this very procedure -- a binary search would disagree.
FSG_HDH int lm_find(const int32_t *words, int begin, int end, int key_, uint32_t max_vocab)
{
    uint32_t before_it = (uint32_t)begin - 1u, before_v = 0, after_it = (uint32_t)end, after_v = max_vocab;
    const uint32_t key = (uint32_t)key_;
    if (key > after_v) return -1;
    while (after_it - before_it > 1) {
        const uint32_t off = key - before_v, range = after_v - before_v, width = after_it - before_it - 1;
        const uint32_t pivot = before_it + (1u + (uint32_t)(off * width) / (range + 1));
        const uint32_t mid = (uint32_t)words[pivot];
        if (mid < key) { before_it = pivot; before_v = mid; }
        else if (mid > key) { after_it = pivot; after_v = mid; }
        else return (int)pivot;
    }
    return -1;
}

// ngram_tg_score(lmset, w, h1, h2) for DICTIONARY word ids (h = -1: no such history word); not yet >> SENSCR_SHIFT
FSG_HDH int lm_tg_score(const LmArr &L, int w_dict, int h1_dict, int h2_dict)
{
    const int w = w_dict < 0 ? -1 : L.widmap[w_dict];
    int h[2], n_hist = 2;
    h[0] = h1_dict < 0 ? -1 : L.widmap[h1_dict];
    h[1] = h2_dict < 0 ? -1 : L.widmap[h2_dict];
    if (w < 0) return L.log_zero;
    if (n_hist > L.order - 1) n_hist = L.order - 1;
    for (int i = 0; i < n_hist; ++i) if (h[i] < 0) { n_hist = i; break; }
    float score = L.uni_prob[w];
    if (n_hist > 0) {
        const int b = lm_find(L.bg_word, L.uni_next[w], L.uni_next[w + 1], h[0], L.max_vocab2);
        if (n_hist == 2) {                                       // full history of a trigram model: cached backoffs
            const float bc0 = L.uni_bo[h[0]];
            float bc1 = 0.0f;
            const int hb = lm_find(L.bg_word, L.uni_next[h[0]], L.uni_next[h[0] + 1], h[1], L.max_vocab2);
            if (hb >= 0) bc1 = L.bg_bo[hb];
            if (b < 0) { score = FSG_FADD(score, bc0); score = FSG_FADD(score, bc1); }
            else {
                const int t = lm_find(L.tg_word, L.bg_next[b], L.bg_next[b + 1], h[1], L.max_vocab3);
                score = t < 0 ? FSG_FADD(L.bg_prob[b], bc1) : L.tg_prob[t];
            }
        }
        else if (L.order == 2) score = b < 0 ? FSG_FADD(score, L.uni_bo[h[0]]) : L.bg_prob[b];       // bigram model, full history
        else score = b < 0 ? FSG_FADD(score, FSG_FADD(0.0f, L.uni_bo[h[0]])) : L.bg_prob[b];          // trigram model, one history word
    }
    const int raw = (int)score;
    return (int)FSG_FADD(FSG_FMUL((float)raw, L.lw), (float)L.log_wip);
}
