// psb_ngf_host.h -- host-side preparation of the n-gram second pass (see psb_ngs_host.h): one channel
// index space over the flat word HMMs (root, word-internal phones, right-context fan-out) of every
// word the second pass can use, per-channel senones and transition matrices.
#pragma once
#include "psb_ngs_host.h"
#include "psb_ngf_core.h"

struct NgfFlat {
    std::vector<int32_t> buf;
    NgfGraph G;
    size_t o_words, o_rs_n, o_rs_cimap, o_ldiph, o_lm, o_inlm, o_pron_off, o_pron_ci, o_ch_off, o_n_int, o_tmatid, o_senid, o_root_ssid, o_sp_index, o_lma;
    int32_t lma_hdr[10];
};

static inline int
ngf_flatten(const int32_t *info, const int32_t *model, long long model_len, const int32_t *lm_arrays, long long lm_arrays_len,
            const int32_t *ci_tmat, const int32_t *ci_ssid, const uint16_t *sseq, int n_sseq, int n_emit, int n_tmat, int n_sen, NgfFlat &o,
            std::string &err)
{
    NgfGraph &G = o.G;
    memset(&G, 0, sizeof(G));
    const int n_words = info[1], n_root = info[2], n_nonroot = info[3], n_1ph = info[4], n_ci = info[6], n_lm = info[26], n_pron = info[33];
    G.use_lma = lm_arrays != nullptr;
    if (n_words <= 0 || n_ci <= 0 || n_lm < 0 || (n_lm == 0 && !G.use_lma) || n_pron <= 0) NGS_FAIL("ngram search: empty tables");
    if (G.use_lma && lm_arr_check(lm_arrays, lm_arrays_len, n_words, err) != 0) return -1;
    G.n_words = n_words; G.n_1ph = n_1ph; G.n_ci = n_ci; G.sil = info[7]; G.n_lm = n_lm; G.n_emit = n_emit;
    G.beam = info[8]; G.fwdflatbeam = info[28]; G.fwdflatwbeam = info[29]; G.min_ef_width = info[30]; G.max_sf_win = info[31];
    memcpy(&G.lwf, &info[32], 4);
    G.pip = info[16]; G.silpen = info[17]; G.fillpen = info[18]; G.start_wid = info[19]; G.finish_wid = info[20];
    G.silence_wid = info[21]; G.filler_start = info[22]; G.filler_end = info[23];
    const size_t nc = (size_t)n_ci;
    if (n_root < 0 || n_nonroot < 0 || n_1ph < 0 || n_ci > 256 || n_lm > 512) NGS_FAIL("ngram search: sizes out of range (at most 256 phones, 512 LM words with the dense tables)");
    {
        const unsigned long long need = (unsigned long long)n_root * 5 + (unsigned long long)n_nonroot * 6 + (unsigned long long)n_words * 8 +
            (unsigned long long)n_1ph * 5 + nc * nc + 3ull * nc * nc * nc + (unsigned long long)n_lm * (n_lm + 1) * (n_lm + 1) +
            (unsigned long long)n_words * 2 + 1 + 2ull * n_pron;
        if (model_len < 0 || (unsigned long long)model_len < need) NGS_FAIL("ngram search: model block holds %lld words, the second pass needs %llu", model_len, need);
    }
    const int32_t *m = model;
    m += (size_t)n_root * 5 + (size_t)n_nonroot * 6;
    const int32_t *words = m; m += (size_t)n_words * 8;
    m += n_1ph + (size_t)n_1ph * 4;
    const int32_t *rs_n = m; m += nc * nc;
    const int32_t *rs_ssid = m; m += nc * nc * nc;
    const int32_t *rs_cimap = m; m += nc * nc * nc;
    const int32_t *ldiph = m; m += nc * nc * nc;
    const int32_t *lm = m; const size_t n_lmtab = (size_t)n_lm * (n_lm + 1) * (n_lm + 1); m += n_lmtab;
    const int32_t *inlm = m; m += n_words;
    const int32_t *pron_off = m; m += n_words + 1;
    const int32_t *pron_ci = m; m += n_pron;
    const int32_t *pron_ssid = m;
    auto wid_ok = [&](int w) { return w >= 0 && w < n_words; };
    if (!wid_ok(G.start_wid) || !wid_ok(G.finish_wid) || !wid_ok(G.silence_wid)) NGS_FAIL("ngram search: special word ids out of range");
    if (pron_off[0] != 0 || pron_off[n_words] != n_pron) NGS_FAIL("pronunciation offsets inconsistent");
    std::vector<int32_t> ch_off((size_t)n_words + 1, 0), n_int((size_t)n_words, 0), root_ssid((size_t)n_words, 0);
    for (int w = 0; w < n_words; ++w) {
        const int32_t *r = words + (size_t)w * 8;
        const int len = pron_off[w + 1] - pron_off[w];
        if (len <= 0 || r[0] < 0 || r[0] >= n_ci || r[1] < 0 || r[1] >= n_ci || r[2] < -1 || r[2] >= n_ci) NGS_FAIL("word %d: pronunciation out of range", w);
        if (r[5] < 0 || r[5] >= n_words || r[7] < -1 || (!G.use_lma && r[7] >= n_lm)) NGS_FAIL("word %d: id out of range", w);
        if ((len == 1) != (r[3] != 0)) NGS_FAIL("word %d: single-phone flag disagrees with its pronunciation", w);
        int n = 0;
        if (r[3]) n = 1;
        else if (inlm[w]) {
            const int nrc = rs_n[(size_t)r[1] * nc + r[2]];
            if (nrc <= 0 || nrc > n_ci) NGS_FAIL("word %d: right-context fan-out %d", w, nrc);
            n = 1 + (len - 2) + nrc;
            n_int[(size_t)w] = len - 2;
        }
        ch_off[(size_t)w + 1] = ch_off[(size_t)w] + n;
        if (ci_ssid[r[0]] < 0 || ci_ssid[r[0]] >= n_sseq) NGS_FAIL("CI phone %d: ssid out of range", r[0]);
        root_ssid[(size_t)w] = ci_ssid[r[0]];
    }
    for (int w = 0; w < n_words; ++w)
        if (!G.use_lma && words[(size_t)w * 8 + 7] < 0 && words[(size_t)words[(size_t)w * 8 + 5] * 8 + 7] < 0) NGS_FAIL("word %d: base word has no LM index", w);
    if (!words[(size_t)G.start_wid * 8 + 3] || !words[(size_t)G.silence_wid * 8 + 3]) NGS_FAIL("<s> / <sil> must be single-phone words");
    const int M_static = ch_off[(size_t)n_words];
    std::vector<int32_t> sp_index((size_t)n_words, -1);
    {
        // state area: the single-phone words' fixed slots + room for an utterance's vocabulary (everything when that
        // is affordable, else PSB_NGF_CHANNELS, default 65536: a vocabulary that does not fit is reported as an error)
        G.n_sp = 0;
        for (int w = 0; w < n_words; ++w) if (words[(size_t)w * 8 + 3]) sp_index[(size_t)w] = G.n_sp++;
        int cap = 65536;
        if (const char *e = getenv("PSB_NGF_CHANNELS")) { const int v = atoi(e); if (v > 0) cap = v; }
        G.M = M_static < cap ? M_static : cap;
        if (G.M < G.n_sp + 1) G.M = G.n_sp + 1;
    }
    G.LW = n_words + 2;
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
    for (int w = 0; w < n_words; ++w) {
        const int32_t *r = words + (size_t)w * 8;
        const int c0 = ch_off[(size_t)w], c1 = ch_off[(size_t)w + 1], ni = n_int[(size_t)w];
        if (c1 == c0) continue;
        for (int q = 0; q < n_ci; ++q) if (ci_tmat[q] < 0 || ci_tmat[q] >= n_tmat) NGS_FAIL("CI phone %d: tmatid out of range", q);
        tmatid[(size_t)c0] = ci_tmat[r[0]];
        for (int k = 0; k < ni; ++k) {
            const int p = pron_off[w] + 1 + k;
            if (pron_ci[p] < 0 || pron_ci[p] >= n_ci) NGS_FAIL("word %d: phone out of range", w);
            tmatid[(size_t)c0 + 1 + k] = ci_tmat[pron_ci[p]];
            if (set_sen(c0 + 1 + k, pron_ssid[p]) != 0) NGS_FAIL("word %d: internal ssid out of range", w);
        }
        for (int c = c0 + 1 + ni; c < c1; ++c) {
            tmatid[(size_t)c] = ci_tmat[r[1]];
            if (set_sen(c, rs_ssid[((size_t)r[1] * nc + r[2]) * nc + (c - (c0 + 1 + ni))]) != 0) NGS_FAIL("word %d: fan-out ssid out of range", w);
        }
        if (!r[3] && (pron_ci[pron_off[w] + 1] < 0 || pron_ci[pron_off[w] + 1] >= n_ci)) NGS_FAIL("word %d: second phone out of range", w);
    }
    for (size_t i = 0; i < nc * nc * nc; ++i) {
        if (rs_cimap[i] < -1 || rs_cimap[i] >= n_ci) NGS_FAIL("rssid cimap out of range");
        if (ldiph[i] != NGS_BAD_SSID && (ldiph[i] < -1 || ldiph[i] >= n_sseq)) NGS_FAIL("ldiph_lc ssid out of range");
    }
    std::vector<int32_t> &b = o.buf;
    b.clear();
    auto put = [&](const int32_t *p, size_t n) { size_t at = b.size(); b.insert(b.end(), p, p + n); return at; };
    o.o_words = put(words, (size_t)n_words * 8); o.o_rs_n = put(rs_n, nc * nc); o.o_rs_cimap = put(rs_cimap, nc * nc * nc);
    o.o_ldiph = put(ldiph, nc * nc * nc); o.o_lm = put(lm, n_lmtab); o.o_inlm = put(inlm, n_words);
    o.o_pron_off = put(pron_off, (size_t)n_words + 1); o.o_pron_ci = put(pron_ci, n_pron);
    o.o_ch_off = put(ch_off.data(), ch_off.size()); o.o_n_int = put(n_int.data(), n_int.size());
    o.o_tmatid = put(tmatid.data(), tmatid.size()); o.o_senid = put(senid.data(), senid.size());
    o.o_root_ssid = put(root_ssid.data(), root_ssid.size());
    o.o_sp_index = put(sp_index.data(), sp_index.size());
    o.o_lma = b.size();
    if (G.use_lma) { put(lm_arrays, (size_t)lm_arrays_len); memcpy(o.lma_hdr, lm_arrays, sizeof(o.lma_hdr)); }
    b.push_back(0);
    return 0;
}

static inline void
ngf_bind(NgfFlat &o, const int32_t *base)
{
    NgfGraph &G = o.G;
    if (G.use_lma) lm_arr_bind(G.lma, o.lma_hdr, base + o.o_lma);
    G.words = base + o.o_words; G.rs_n = base + o.o_rs_n; G.rs_cimap = base + o.o_rs_cimap; G.ldiph = base + o.o_ldiph;
    G.lm = base + o.o_lm; G.inlm = base + o.o_inlm; G.pron_off = base + o.o_pron_off; G.pron_ci = base + o.o_pron_ci;
    G.ch_off = base + o.o_ch_off; G.n_int = base + o.o_n_int; G.tmatid = base + o.o_tmatid; G.senid = base + o.o_senid;
    G.root_ssid = base + o.o_root_ssid; G.sp_index = base + o.o_sp_index;
}
