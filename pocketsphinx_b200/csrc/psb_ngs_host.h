// psb_ngs_host.h -- host-side preparation of the n-gram first pass: derives, from the flattened
// search a caller hands over (the sections oracle/ref_driver.c:refdrv_fwdtree documents), the tables
// the phase code needs (one channel index space, fan-out slots, parents, per-channel senones) and
// lays everything out as one int32 block.  Shared by the kernel's ABI and tests/emul/ngs_emul.cpp.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "psb_ngs_core.h"

struct NgsFlat {
    std::vector<int32_t> buf;
    NgsGraph G;                 // pointers are offsets into buf until ngs_bind()
    size_t o_roots, o_nonroot, o_words, o_w1ph, o_r1ph, o_rs_n, o_rs_ssid, o_rs_cimap, o_ldiph, o_lm, o_wc_off, o_w2h1,
           o_parent, o_tmatid, o_senid, o_lma;
    int32_t lma_hdr[10];
};

#define NGS_FAIL(...) do { char b_[256]; snprintf(b_, sizeof b_, __VA_ARGS__); err = b_; return -1; } while (0)

// The optional array LM (psb_lm_core.h): bounds, so that the device's searches cannot leave the block
// (ranges need not be sorted: shipped models contain unsorted ones and the search is the reference's own).
static inline int
lm_arr_check(const int32_t *a, long long len, int n_words, std::string &err)
{
    if (len < 10) NGS_FAIL("LM arrays: block too short");
    const long long order = a[0], V = a[1], n2 = a[2], n3 = a[3];
    if (order < 1 || order > 3 || V <= 0 || n2 < 0 || n3 < 0 || a[7] != n_words) NGS_FAIL("LM arrays: bad header (order %lld, %lld unigrams, %d words)", order, V, a[7]);
    if ((long long)lm_arr_words(a) != len) NGS_FAIL("LM arrays: block holds %lld words, its header needs %zu", len, lm_arr_words(a));
    if ((order < 2 && n2 != 0) || (order < 3 && n3 != 0)) NGS_FAIL("LM arrays: n-gram counts do not fit the order");
    LmArr L;
    lm_arr_bind(L, a, a);
    for (int w = 0; w < n_words; ++w) if (L.widmap[w] < -1 || L.widmap[w] >= V) NGS_FAIL("LM arrays: widmap[%d] out of range", w);
    if (L.uni_next[0] != 0 || L.uni_next[V] < 0 || L.uni_next[V] > n2) NGS_FAIL("LM arrays: unigram ranges leave the bigram array");
    const long long n2_used = L.uni_next[V];            /* (the header may count a few more than the trie links) */
    for (long long w = 0; w < V; ++w) {
        if (L.uni_next[w + 1] < L.uni_next[w]) NGS_FAIL("LM arrays: unigram ranges not monotone");
        for (int p = L.uni_next[w]; p < L.uni_next[w + 1]; ++p)
            if (L.bg_word[p] < 0 || L.bg_word[p] >= V) NGS_FAIL("LM arrays: bigrams of word %lld out of range", w);
    }
    if (n2_used > 0 && (L.bg_next[0] != 0 || L.bg_next[n2_used] < 0 || L.bg_next[n2_used] > n3)) NGS_FAIL("LM arrays: bigram ranges leave the trigram array");
    for (long long b = 0; b < n2_used; ++b) {
        if (L.bg_next[b + 1] < L.bg_next[b]) NGS_FAIL("LM arrays: bigram ranges not monotone");
        for (int p = L.bg_next[b]; p < L.bg_next[b + 1]; ++p)
            if (L.tg_word[p] < 0 || L.tg_word[p] >= V) NGS_FAIL("LM arrays: trigrams of bigram %lld out of range", b);
    }
    return 0;
}

// info[40] + model sections as exported; ci_tmat[n_ci]; sseq [n_sseq][n_emit].
static inline int
ngs_flatten(const int32_t *info, const int32_t *model, long long model_len, const int32_t *lm_arrays, long long lm_arrays_len,
            const int32_t *ci_tmat, const uint16_t *sseq, int n_sseq, int n_emit, int n_tmat, int n_sen, NgsFlat &o, std::string &err)
{
    NgsGraph &G = o.G;
    memset(&G, 0, sizeof(G));
    G.n_words = info[1]; G.n_root = info[2]; G.n_nonroot = info[3]; G.n_1ph = info[4]; G.n_1ph_lm = info[5];
    G.n_ci = info[6]; G.sil = info[7]; G.beam = info[8]; G.pbeam = info[9]; G.wbeam = info[10]; G.lpbeam = info[11];
    G.lponlybeam = info[12]; G.maxhmmpf = info[13]; G.maxwpf = info[14]; G.nwpen = info[15]; G.pip = info[16];
    G.silpen = info[17]; G.fillpen = info[18]; G.start_wid = info[19]; G.finish_wid = info[20]; G.silence_wid = info[21];
    G.filler_start = info[22]; G.filler_end = info[23]; G.n_lm = info[26]; G.n_emit = n_emit;
    G.use_lma = lm_arrays != nullptr;
    if (G.n_words <= 0 || G.n_root < 0 || G.n_nonroot < 0 || G.n_1ph <= 0 || G.n_ci <= 0 || G.n_lm < 0 || (G.n_lm == 0 && !G.use_lma)) NGS_FAIL("ngram search: empty tables");
    if (G.use_lma && lm_arr_check(lm_arrays, lm_arrays_len, G.n_words, err) != 0) return -1;
    if (G.n_1ph_lm > G.n_1ph) NGS_FAIL("ngram search: n_1ph_LMwords > n_1ph_words");
    const size_t nc = (size_t)G.n_ci;
    if (G.n_ci > 256 || G.n_lm > 512) NGS_FAIL("ngram search: %d phones / %d LM words exceed what the dense tables are meant for (256 / 512)", G.n_ci, G.n_lm);
    {
        const unsigned long long need = (unsigned long long)G.n_root * 5 + (unsigned long long)G.n_nonroot * 6 + (unsigned long long)G.n_words * 8 +
            (unsigned long long)G.n_1ph * 5 + nc * nc + 3ull * nc * nc * nc + (unsigned long long)G.n_lm * (G.n_lm + 1) * (G.n_lm + 1);
        if (model_len < 0 || (unsigned long long)model_len < need) NGS_FAIL("ngram search: model block holds %lld words, the sizes in info need %llu", model_len, need);
    }
    const int32_t *m = model;
    const int32_t *roots = m; m += (size_t)G.n_root * 5;
    const int32_t *nonroot = m; m += (size_t)G.n_nonroot * 6;
    const int32_t *words = m; m += (size_t)G.n_words * 8;
    const int32_t *w1ph = m; m += G.n_1ph;
    const int32_t *r1ph = m; m += (size_t)G.n_1ph * 4;
    const int32_t *rs_n = m; m += nc * nc;
    const int32_t *rs_ssid = m; m += nc * nc * nc;
    const int32_t *rs_cimap = m; m += nc * nc * nc;
    const int32_t *ldiph = m; m += nc * nc * nc;
    const int32_t *lm = m;
    const size_t n_lmtab = (size_t)G.n_lm * (G.n_lm + 1) * (G.n_lm + 1);
    auto wid_ok = [&](int w) { return w >= 0 && w < G.n_words; };
    if (!wid_ok(G.start_wid) || !wid_ok(G.finish_wid) || !wid_ok(G.silence_wid)) NGS_FAIL("ngram search: special word ids out of range");
    std::vector<int32_t> wc_off((size_t)G.n_words + 1, 0), w2h1((size_t)G.n_words, -1), parent((size_t)G.n_nonroot, INT32_MIN);
    for (int w = 0; w < G.n_words; ++w) {
        const int32_t *r = words + (size_t)w * 8;
        if (r[0] < 0 || r[0] >= G.n_ci || r[1] < 0 || r[1] >= G.n_ci || r[2] < -1 || r[2] >= G.n_ci) NGS_FAIL("word %d: phone out of range", w);
        if (r[6] < -1 || r[6] >= G.n_words || r[5] < 0 || r[5] >= G.n_words || r[7] < -1 || (!G.use_lma && r[7] >= G.n_lm)) NGS_FAIL("word %d: id out of range", w);
        int n = 0;
        if (!r[3]) {
            if (r[2] < 0) NGS_FAIL("word %d: multi-phone word without a second-last phone", w);
            n = rs_n[(size_t)r[1] * nc + r[2]];
            if (n < 0 || n > G.n_ci) NGS_FAIL("word %d: right-context fan-out %d", w, n);
        }
        wc_off[(size_t)w + 1] = wc_off[(size_t)w] + n;
    }
    for (int w = 0; w < G.n_words; ++w)
        if (!G.use_lma && words[(size_t)w * 8 + 7] < 0 && words[(size_t)words[(size_t)w * 8 + 5] * 8 + 7] < 0) NGS_FAIL("word %d: base word has no LM index", w);
    G.n_rcchan = wc_off[(size_t)G.n_words];
    for (int i = 0; i < G.n_1ph; ++i) {
        if (!wid_ok(w1ph[i]) || !words[(size_t)w1ph[i] * 8 + 3]) NGS_FAIL("single-phone word list entry %d is not a single-phone word", i);
        w2h1[(size_t)w1ph[i]] = i;
    }
    if (w2h1[(size_t)G.start_wid] < 0 || w2h1[(size_t)G.silence_wid] < 0) NGS_FAIL("<s> / <sil> must be single-phone words");
    long visited = 0;
    auto chain = [&](int first, int par) -> int {
        for (int c = first; c >= 0; c = nonroot[(size_t)c * 6 + 5]) {
            if (c >= G.n_nonroot || ++visited > G.n_nonroot || parent[(size_t)c] != INT32_MIN) return -1;
            parent[(size_t)c] = par;
        }
        return 0;
    };
    for (int i = 0; i < G.n_root; ++i) {
        const int32_t *r = roots + (size_t)i * 5;
        if (r[0] < 0 || r[0] >= G.n_ci || r[1] < 0 || r[1] >= G.n_ci || r[2] < -1 || r[2] >= G.n_words || r[3] < -1 || r[4] < 0 || r[4] >= n_tmat)
            NGS_FAIL("root channel %d out of range", i);
        if (chain(r[3], -i - 1) != 0) NGS_FAIL("lextree is not a tree below root %d", i);
    }
    for (int i = 0; i < G.n_nonroot; ++i) {
        const int32_t *r = nonroot + (size_t)i * 6;
        if (r[0] < 0 || r[0] >= n_sseq || r[1] < 0 || r[1] >= n_tmat || r[2] < 0 || r[2] >= G.n_ci || r[3] < -1 || r[3] >= G.n_words || r[4] < -1)
            NGS_FAIL("non-root channel %d out of range", i);
        if (chain(r[4], i) != 0) NGS_FAIL("lextree is not a tree below channel %d", i);
    }
    for (int i = 0; i < G.n_nonroot; ++i) if (parent[(size_t)i] == INT32_MIN) NGS_FAIL("non-root channel %d is unreachable", i);
    G.o_nonroot = G.n_root; G.o_1ph = G.o_nonroot + G.n_nonroot; G.o_rc = G.o_1ph + G.n_1ph;
    {
        // fan-out pool: a block per word that can need one when that is affordable (then it can never run dry),
        // else a fixed number (PSB_NGS_BLOCKS, default 8192: large vocabularies; a dry pool is reported as an error)
        int n_multi = 0;
        G.RB = 1;
        for (int w = 0; w < G.n_words; ++w) {
            const int n = wc_off[(size_t)w + 1] - wc_off[(size_t)w];
            if (n > 0) ++n_multi;
            if (n > G.RB) G.RB = n;
        }
        int cap = 8192;
        if (const char *e = getenv("PSB_NGS_BLOCKS")) { const int v = atoi(e); if (v > 0) cap = v; }
        G.n_blocks = n_multi < cap ? (n_multi > 0 ? n_multi : 1) : cap;
        G.M = G.o_rc + G.n_blocks * G.RB;
    }
    const int M_static = G.o_rc + G.n_rcchan;
    {
        int lw = G.n_root;
        if (G.n_nonroot > lw) lw = G.n_nonroot;
        if (G.n_words + G.n_1ph > lw) lw = G.n_words + G.n_1ph;
        G.LW = lw + 2;
    }
    std::vector<int32_t> tmatid((size_t)M_static, 0), senid((size_t)M_static * n_emit, NGS_BAD_SSID);
    auto set_sen = [&](int c, int ssid) -> int {
        if (ssid < 0 || ssid >= n_sseq) return -1;
        for (int s = 0; s < n_emit; ++s) {
            const int v = sseq[(size_t)ssid * n_emit + s];
            if (v >= n_sen) return -1;
            senid[(size_t)c * n_emit + s] = v;
        }
        return 0;
    };
    for (int i = 0; i < G.n_root; ++i) tmatid[(size_t)i] = roots[(size_t)i * 5 + 4];
    for (int i = 0; i < G.n_nonroot; ++i) {
        tmatid[(size_t)G.o_nonroot + i] = nonroot[(size_t)i * 6 + 1];
        if (set_sen(G.o_nonroot + i, nonroot[(size_t)i * 6]) != 0) NGS_FAIL("non-root channel %d: ssid out of range", i);
    }
    for (int i = 0; i < G.n_1ph; ++i) {
        const int32_t *r = r1ph + (size_t)i * 4;
        if (r[0] < 0 || r[0] >= G.n_ci || r[1] < 0 || r[1] >= G.n_ci || r[2] < 0 || r[2] >= n_sseq || r[3] < 0 || r[3] >= n_tmat)
            NGS_FAIL("single-phone channel %d out of range", i);
        tmatid[(size_t)G.o_1ph + i] = r[3];
    }
    for (int w = 0; w < G.n_words; ++w) {
        const int32_t *r = words + (size_t)w * 8;
        for (int k = wc_off[(size_t)w]; k < wc_off[(size_t)w + 1]; ++k) {
            const int rc = k - wc_off[(size_t)w];
            if (ci_tmat[r[1]] < 0 || ci_tmat[r[1]] >= n_tmat) NGS_FAIL("CI phone %d: tmatid out of range", r[1]);
            tmatid[(size_t)G.o_rc + k] = ci_tmat[r[1]];
            if (set_sen(G.o_rc + k, rs_ssid[((size_t)r[1] * nc + r[2]) * nc + rc]) != 0) NGS_FAIL("word %d: fan-out ssid out of range", w);
        }
    }
    for (size_t i = 0; i < nc * nc * nc; ++i) {
        if (rs_cimap[i] < -1 || rs_cimap[i] >= G.n_ci) NGS_FAIL("rssid cimap out of range");
        if (ldiph[i] != NGS_BAD_SSID && (ldiph[i] < -1 || ldiph[i] >= n_sseq)) NGS_FAIL("ldiph_lc ssid out of range");
    }
    std::vector<int32_t> &b = o.buf;
    b.clear();
    auto put = [&](const int32_t *p, size_t n) { size_t at = b.size(); b.insert(b.end(), p, p + n); return at; };
    o.o_roots = put(roots, (size_t)G.n_root * 5); o.o_nonroot = put(nonroot, (size_t)G.n_nonroot * 6);
    o.o_words = put(words, (size_t)G.n_words * 8); o.o_w1ph = put(w1ph, G.n_1ph); o.o_r1ph = put(r1ph, (size_t)G.n_1ph * 4);
    o.o_rs_n = put(rs_n, nc * nc); o.o_rs_ssid = put(rs_ssid, nc * nc * nc); o.o_rs_cimap = put(rs_cimap, nc * nc * nc);
    o.o_ldiph = put(ldiph, nc * nc * nc); o.o_lm = put(lm, n_lmtab);
    o.o_wc_off = put(wc_off.data(), wc_off.size()); o.o_w2h1 = put(w2h1.data(), w2h1.size());
    o.o_parent = put(parent.data(), parent.size()); o.o_tmatid = put(tmatid.data(), tmatid.size());
    o.o_senid = put(senid.data(), senid.size());
    o.o_lma = b.size();
    if (G.use_lma) { put(lm_arrays, (size_t)lm_arrays_len); memcpy(o.lma_hdr, lm_arrays, sizeof(o.lma_hdr)); }
    b.push_back(0);
    return 0;
}

static inline void
ngs_bind(NgsFlat &o, const int32_t *base)
{
    NgsGraph &G = o.G;
    if (G.use_lma) lm_arr_bind(G.lma, o.lma_hdr, base + o.o_lma);
    G.roots = base + o.o_roots; G.nonroot = base + o.o_nonroot; G.words = base + o.o_words; G.w1ph = base + o.o_w1ph;
    G.r1ph = base + o.o_r1ph; G.rs_n = base + o.o_rs_n; G.rs_ssid = base + o.o_rs_ssid; G.rs_cimap = base + o.o_rs_cimap;
    G.ldiph = base + o.o_ldiph; G.lm = base + o.o_lm; G.wc_off = base + o.o_wc_off; G.w2h1 = base + o.o_w2h1;
    G.parent = base + o.o_parent; G.tmatid = base + o.o_tmatid; G.senid = base + o.o_senid;
}
