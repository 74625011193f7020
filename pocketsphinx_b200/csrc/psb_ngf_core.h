// psb_ngf_core.h -- the second pass of n-gram decoding (ngram_search_fwdflat.c) for ONE utterance
// as block-wide data-parallel phases; conventions of psb_fsg_core.h / psb_ngs_core.h (the same
// source is compiled into ngs_fwdflat_kernel and into the host harness tests/emul/ngf_emul.cpp).
//
// The second pass is word-local: every word of the utterance vocabulary is a flat chain
// root -> word-internal phones -> right-context fan-out, and evaluation, pruning, phone transitions
// and ALL exits of a word touch only that word (fwdflat_prune_chan :483-607).  So: one thread per
// active word runs the reference's chain logic as it stands; what crosses words is
//  * the utterance vocabulary (build_fwdflat_wordlist :224-300): per (start frame, word) the first
//    and last end frame in backpointer-table order -- the table's entries bucketed by word, grouped
//    by start frame inside the word; a word's place in the list is (first start frame with a
//    surviving node, later-created nodes first), a rank;
//  * the order of the backpointer table: one entry per exiting word, in active-list order = scan;
//  * word transitions (:643-782): per target word the first maximum over the frame's exits of
//    exit score + float-scaled trigram score; targets are the words with a node whose start frame
//    lies in [frame - win, frame + win): the word's sorted node list answers that;
//  * the next active list (:852-866): vocabulary order, then ids >= <s> ascending = two scans.
#pragma once
#include "psb_ngs_core.h"

struct NgfGraph {
    int n_words, n_1ph, n_ci, sil, n_lm, n_emit;
    int beam, fwdflatbeam, fwdflatwbeam, min_ef_width, max_sf_win, pip, silpen, fillpen;
    int start_wid, finish_wid, silence_wid, filler_start, filler_end;
    int M, LW;                 // M: channels with per-utterance state (single-phone words + the utterance vocabulary's chains)
    int n_sp;                  // single-phone words (fixed state slots 0 .. n_sp)
    float lwf;
    const int32_t *words;      // [n_words][8]  first, last, last2, single, filler, basewid, homophone, lmidx
    const int32_t *rs_n, *rs_cimap, *ldiph, *lm, *inlm;
    const int32_t *pron_off, *pron_ci;
    const int32_t *ch_off;     // [n_words+1]  STATIC index (tmatid, senid) of a word's channels: root, internal phones, fan-out
                               //              (single-phone: root only); state lives at W.wbase[w] + position
    const int32_t *sp_index;   // [n_words]    slot of a single-phone word among the single-phone words, or -1
    const int32_t *n_int;      // [n_words]    word-internal phones (pronlen - 2), 0 for single-phone words
    const int32_t *tmatid;     // [M]
    const int32_t *senid;      // [M][n_emit]  non-multiplexed channels
    const int32_t *root_ssid;  // [n_words]    senone sequence a root starts from (its CI phone's)
    int use_lma;               // 1: trigram scores from the sorted-array LM below instead of the dense table `lm`
    LmArr lma;
};

struct NgfWork {
    fsg_wp score, hist, mss;                     // [n_emit][M]
    fsg_wp out_score, out_hist, best, frame;    // [M]
    fsg_wp awl[2];                                 // [n_words]
    fsg_wp word_active, wordlist, first_sf, wl_key;   // [n_words]
    fsg_wp wbase;                                  // [n_words] first state channel of a word in this utterance, or -1
    fsg_wp wstart, wn;                             // [n_words+1], [n_words]: a word's segment in the node arrays, surviving nodes in it
    fsg_wp nd_ent, nd_sf, nd_key;                  // [n_bp_in]: per word, its first-pass entries, then its surviving nodes' start frames
                                                   //            (ascending) and creation keys (index of the node's first entry)
    fsg_wp cnt, cnt2, cnt3;                      // [LW]
    fsg_wp bp, bss, bp_idx;                      // outputs
    const int32_t *bp_in;                            // first pass: [n_bp_in][10]
    int n_bp_in, bp_cap, bss_cap, T;
};

struct NgfScalars {
    fsg_int cur, n_awl, n_awl_nxt, nwd, best, best_score, bpidx, bss_head, stop, error, n_done, renorm, norm;
    fsg_int silrc_score, silrc_bp;
    int scan[34];
};

FSG_HDH size_t ngf_work_words(const NgfGraph &G, int T, int n_bp_cap)
{
    const size_t M = (size_t)G.M, N = (size_t)G.n_emit, nw = (size_t)G.n_words;
    (void)T;
    return 3 * N * M + 4 * M + 2 * (nw + 1) + 5 * nw + (nw + 1) + nw + 3 * ((size_t)n_bp_cap + 1) + 3 * (size_t)G.LW;
}

FSG_HD void ngf_work_carve(int32_t *b, const NgfGraph &G, int T, int n_bp_cap, NgfWork &W)
{
    const size_t M = (size_t)G.M, N = (size_t)G.n_emit, nw = (size_t)G.n_words;
    W.score = b; b += N * M;  W.hist = b; b += N * M;  W.mss = b; b += N * M;
    W.out_score = b; b += M;  W.out_hist = b; b += M;  W.best = b; b += M;  W.frame = b; b += M;
    W.awl[0] = b; b += nw + 1;  W.awl[1] = b; b += nw + 1;
    W.word_active = b; b += nw;  W.wordlist = b; b += nw;  W.first_sf = b; b += nw;  W.wl_key = b; b += nw;  W.wbase = b; b += nw;
    W.wstart = b; b += nw + 1;  W.wn = b; b += nw;
    W.nd_ent = b; b += (size_t)n_bp_cap + 1;  W.nd_sf = b; b += (size_t)n_bp_cap + 1;  W.nd_key = b; b += (size_t)n_bp_cap + 1;
    W.cnt = b; b += G.LW;  W.cnt2 = b; b += G.LW;  W.cnt3 = b;
    W.T = T;
}

FSG_HD void ngf_clear(const NgfGraph &G, const NgfWork &W, int c)
{
    for (int s = 0; s < G.n_emit; ++s) { W.score[s * G.M + c] = FSG_WORST_SCORE; W.hist[s * G.M + c] = -1; }
    W.out_score[c] = FSG_WORST_SCORE; W.out_hist[c] = -1; W.best[c] = FSG_WORST_SCORE; W.frame[c] = -1;
}

FSG_HD void ngf_normalize(const NgfGraph &G, const NgfWork &W, int c, int norm)  /* hmm_normalize, hmm.c:206-217 */
{
    for (int s = 0; s < G.n_emit; ++s) {
        const int v = W.score[s * G.M + c];
        if (v > FSG_WORST_SCORE) W.score[s * G.M + c] = v - norm;
    }
    const int o = W.out_score[c];
    if (o > FSG_WORST_SCORE) W.out_score[c] = o - norm;
}

FSG_HD void ngf_clear_scores(const NgfGraph &G, const NgfWork &W, int c)      /* hmm_clear_scores, hmm.c:167-178 */
{
    for (int s = 0; s < G.n_emit; ++s) W.score[s * G.M + c] = FSG_WORST_SCORE;
    W.out_score[c] = FSG_WORST_SCORE; W.best[c] = FSG_WORST_SCORE;
}

FSG_HD void ngf_enter(const NgfWork &W, int c, int score, int hist, int nf) { W.score[c] = score; W.hist[c] = hist; W.frame[c] = nf; }

FSG_HD int ngf_tg(const NgfGraph &G, int w, int h1, int h2)
{
    if (G.use_lma) return lm_tg_score(G.lma, w, h1, h2) >> 10;           /* >> SENSCR_SHIFT */
    const int n = G.n_lm + 1;
    const int a = NGS_W(G, w, 7), b = h1 < 0 ? 0 : NGS_W(G, h1, 7) + 1, c = h2 < 0 ? 0 : NGS_W(G, h2, 7) + 1;
    return G.lm[((size_t)a * n + b) * n + c];
}

// ngram_fwdflat_start :371-414
FSG_HD void ngf_start(const NgfGraph &G, const NgfWork &W, NgfScalars *S)
{
    const int nw = G.n_words, T = W.T;
    FSG_FOR(w, nw) { W.word_active[w] = 0; W.first_sf[w] = -1; W.wbase[w] = G.sp_index[w]; }
    FSG_IF_LEADER {
        S->cur = 0; S->n_awl = 0; S->n_awl_nxt = 0; S->nwd = 0; S->best_score = 0; S->bpidx = 0; S->bss_head = 0;
        S->stop = 0; S->error = 0; S->n_done = 0;
    }
    FSG_SYNC();
    if (W.n_bp_in < 0) {
        // no first pass (-fwdtree no): ngram_fwdflat_expand_all :61-87 -- every LM word is in the
        // vocabulary (in id order) and may follow every exit (get_expand_wordlist :615-618)
        FSG_FOR(w, nw) {
            const int in = G.inlm[w] ? 1 : 0;
            W.wn[w] = 0; W.wstart[w] = 0;
            W.first_sf[w] = in ? 0 : -1;
            W.wl_key[w] = -w;
        }
        FSG_SYNC();
    }
    else {
        // build_fwdflat_wordlist :224-300.  The reference keeps, per start frame, a list of (word, first end
        // frame, last end frame) nodes, created and updated in backpointer-table order.  Here: the table's
        // entries are bucketed by word (count, scan, fill), and each word's thread -- its entries are few --
        // sorts them back into table order, groups them by start frame, drops the nodes with too few end
        // points (or </s> not ending the utterance) and leaves its surviving nodes sorted by start frame.
        const int n_in = W.n_bp_in;
        FSG_FOR(w, nw) { W.wn[w] = 0; W.cnt[w] = 0; }
        FSG_SYNC();
        FSG_FOR(i, n_in) {
            const int32_t *b = W.bp_in + (size_t)i * NGS_BP_ROW;
            const int sf = b[3] < 0 ? 0 : W.bp_in[(size_t)b[3] * NGS_BP_ROW] + 1, wid = b[2];
            if (G.inlm[wid] && sf < T) FSG_ATOMIC_ADD_AT(W.cnt, wid, 1);
        }
        FSG_SYNC();
        FSG_FOR(w, nw) W.wstart[w] = W.cnt[w];
        FSG_SYNC();
        const int n_used = fsg_exscan(W.wstart, nw, S->scan);
        FSG_IF_LEADER W.wstart[nw] = n_used;
        FSG_FOR(w, nw) W.cnt[w] = 0;
        FSG_SYNC();
        FSG_FOR(i, n_in) {
            const int32_t *b = W.bp_in + (size_t)i * NGS_BP_ROW;
            const int sf = b[3] < 0 ? 0 : W.bp_in[(size_t)b[3] * NGS_BP_ROW] + 1, wid = b[2];
            if (G.inlm[wid] && sf < T) W.nd_ent[W.wstart[wid] + FSG_ATOMIC_FETCH_ADD_AT(W.cnt, wid, 1)] = i;
        }
        FSG_SYNC();
        FSG_FOR(w, nw) {
            const int s0 = W.wstart[w], n = W.wstart[w + 1] - s0;
            int nn = 0, f0 = -1, k0 = -1;
            for (int a = 1; a < n; ++a) {                                     // back into table order
                const int v = W.nd_ent[s0 + a];
                int p = a - 1;
                while (p >= 0 && W.nd_ent[s0 + p] > v) { W.nd_ent[s0 + p + 1] = W.nd_ent[s0 + p]; --p; }
                W.nd_ent[s0 + p + 1] = v;
            }
            for (int a = 0; a < n; ++a) {                                     // group by start frame, in creation order
                const int i = W.nd_ent[s0 + a];
                const int32_t *b = W.bp_in + (size_t)i * NGS_BP_ROW;
                const int sf = b[3] < 0 ? 0 : W.bp_in[(size_t)b[3] * NGS_BP_ROW] + 1;
                int q = 0;
                while (q < nn && W.nd_sf[s0 + q] != sf) ++q;
                if (q == nn) { W.nd_sf[s0 + nn] = sf; W.nd_key[s0 + nn] = i; W.nd_ent[s0 + nn] = i; ++nn; }   // nd_ent reused: last entry of node q
                else W.nd_ent[s0 + q] = i;                                    // (q <= a always: slot q is no longer needed as input)
            }
            int m = 0;
            for (int q = 0; q < nn; ++q) {                                    // too few end points / </s> not at the end: drop
                const int fef = W.bp_in[(size_t)W.nd_key[s0 + q] * NGS_BP_ROW], lef = W.bp_in[(size_t)W.nd_ent[s0 + q] * NGS_BP_ROW];
                if (lef - fef < G.min_ef_width || (w == G.finish_wid && lef < T - 1)) continue;
                W.nd_sf[s0 + m] = W.nd_sf[s0 + q]; W.nd_key[s0 + m] = W.nd_key[s0 + q]; ++m;
            }
            for (int a = 1; a < m; ++a) {                                     // by start frame
                const int vs = W.nd_sf[s0 + a], vk = W.nd_key[s0 + a];
                int p = a - 1;
                while (p >= 0 && W.nd_sf[s0 + p] > vs) { W.nd_sf[s0 + p + 1] = W.nd_sf[s0 + p]; W.nd_key[s0 + p + 1] = W.nd_key[s0 + p]; --p; }
                W.nd_sf[s0 + p + 1] = vs; W.nd_key[s0 + p + 1] = vk;
            }
            if (m > 0) { f0 = W.nd_sf[s0]; k0 = W.nd_key[s0]; }
            W.wn[w] = m; W.first_sf[w] = f0; W.wl_key[w] = k0;
        }
        FSG_SYNC();
    }
    // utterance vocabulary: by (first start frame ascending, node creation order descending)
    FSG_FOR(w, nw) {
        const int f0 = W.first_sf[w];
        if (f0 < 0) continue;
        int rank = 0;
        for (int v = 0; v < nw; ++v) {
            const int fv = W.first_sf[v];
            if (v == w || fv < 0) continue;
            if (fv < f0 || (fv == f0 && W.wl_key[v] > W.wl_key[w])) ++rank;
        }
        W.wordlist[rank] = w;
        FSG_ATOMIC_ADD(&S->nwd, 1);
    }
    FSG_SYNC();
    // build_fwdflat_chan :306-368: state channels for the utterance vocabulary only (single-phone words keep
    // their fixed slots), laid out in vocabulary order by a scan over the chain lengths
    const int nwd = S->nwd;
    FSG_FOR(k, nwd) { const int w = W.wordlist[k]; W.cnt[k] = NGS_W(G, w, 3) ? 0 : G.ch_off[w + 1] - G.ch_off[w]; }
    FSG_SYNC();
    const int n_chain = fsg_exscan(W.cnt, nwd, S->scan);
    if (G.n_sp + n_chain > G.M) {
        FSG_IF_LEADER S->error = 3;                                           // vocabulary does not fit the state area
        FSG_SYNC();
        return;
    }
    FSG_FOR(k, nwd) { const int w = W.wordlist[k]; if (!NGS_W(G, w, 3)) W.wbase[w] = G.n_sp + W.cnt[k]; }
    FSG_FOR(c, G.n_sp + n_chain) { ngf_clear(G, W, c); for (int s = 0; s < G.n_emit; ++s) W.mss[s * G.M + c] = NGS_BAD_SSID; }
    FSG_SYNC();
    FSG_FOR(w, nw) if (W.wbase[w] >= 0) W.mss[W.wbase[w]] = G.root_ssid[w];
    FSG_SYNC();
    FSG_IF_LEADER {
        ngf_enter(W, W.wbase[G.start_wid], 0, -1, 0);
        W.awl[0][0] = G.start_wid; S->n_awl = 1;
    }
    FSG_SYNC();
}

template <class Eval>
FSG_HD void ngf_step(const NgfGraph &G, const NgfWork &W, NgfScalars *S, int cf, Eval &eval)
{
    const int nf = cf + 1, cur = S->cur, nxt = cur ^ 1, nw = S->n_awl, pip = G.pip, nwords = G.n_words, T = W.T;
    const fsg_wp awl = W.awl[cur];
    fsg_wp nawl = W.awl[nxt];
    FSG_IF_LEADER {
        W.bp_idx[cf] = S->bpidx;
        if (S->best_score <= FSG_WORST_SCORE) S->stop = 1;
        S->renorm = 0;
        if (!S->stop && S->best_score + 2 * G.beam < FSG_WORST_SCORE) { S->renorm = 1; S->norm = S->best_score; }
        S->best = FSG_WORST_SCORE;
    }
    FSG_SYNC();
    if (S->stop || S->error) return;
    if (S->renorm) {
        // fwdflat_renormalize_scores :785-810: hmm_normalize on the active words' channels of this frame
        const int norm = S->norm;
        FSG_FOR(j, nw) {
            const int w = awl[j], s0 = G.ch_off[w], c0 = W.wbase[w], c1 = c0 + (G.ch_off[w + 1] - s0);
            for (int c = c0; c < c1; ++c)
                if (W.frame[c] == cf) ngf_normalize(G, W, c, norm);
        }
        FSG_SYNC();
    }
    // fwdflat_eval_chan :445-480
    FSG_FOR(j, nw) {
        const int w = awl[j], s0 = G.ch_off[w], c0 = W.wbase[w], c1 = c0 + (G.ch_off[w + 1] - s0);
        int b = FSG_WORST_SCORE;
        if (W.frame[c0] == cf) { const int sc = eval(W, c0, true, s0); if (w != G.finish_wid && sc > b) b = sc; }
        for (int c = c0 + 1; c < c1; ++c)
            if (W.frame[c] == cf) { const int sc = eval(W, c, false, s0 + (c - c0)); if (sc > b) b = sc; }
        FSG_ATOMIC_MAX(&S->best, b);
    }
    FSG_FOR(w, nwords) W.word_active[w] = 0;
    FSG_SYNC();
    FSG_IF_LEADER S->best_score = S->best;
    FSG_SYNC();
    const int thresh = S->best_score + G.fwdflatbeam, wordthresh = S->best_score + G.fwdflatwbeam;
    // fwdflat_prune_chan :483-607, pass 1: which words exit (decided by what evaluation left)
    FSG_FOR(j, nw) {
        const int w = awl[j], c0 = W.wbase[w], c1 = c0 + (G.ch_off[w + 1] - G.ch_off[w]), ni = G.n_int[w];
        int ex = 0;
        if (NGS_W(G, w, 3)) ex = W.frame[c0] == cf && W.best[c0] > thresh && W.out_score[c0] > wordthresh;
        else
            for (int c = c0 + 1 + ni; c < c1; ++c)
                if (W.frame[c] == cf && W.best[c] > thresh && W.out_score[c] > wordthresh) ex = 1;
        W.cnt[j] = ex;
        W.cnt2[j] = (ex && !NGS_W(G, w, 3)) ? c1 - (c0 + 1 + ni) : 0;
    }
    FSG_SYNC();
    const int n_new_bp = fsg_exscan(W.cnt, nw, S->scan);
    const int n_new_bss = fsg_exscan(W.cnt2, nw, S->scan);
    if (S->bpidx + n_new_bp > W.bp_cap || S->bss_head + n_new_bss > W.bss_cap) {
        FSG_IF_LEADER S->error = 1;
        FSG_SYNC();
        return;
    }
    // pass 2: the chain logic, one word per thread
    FSG_FOR(j, nw) {
        const int w = awl[j], c0 = W.wbase[w], c1 = c0 + (G.ch_off[w + 1] - G.ch_off[w]), ni = G.n_int[w], single = NGS_W(G, w, 3);
        const int rc0 = c0 + 1 + ni;
        const int new_bp = S->bpidx + W.cnt[j], new_s = S->bss_head + W.cnt2[j];
        int entry = -1;
        if (W.frame[c0] == cf && W.best[c0] > thresh) {
            int ns = W.out_score[c0];
            W.frame[c0] = nf; W.word_active[w] = 1;
            if (!single) {
                ns += pip;
                if (ns > thresh) {
                    if (ni == 0) {
                        for (int c = rc0; c < c1; ++c) if (W.frame[c] < cf || ns > W.score[c]) ngf_enter(W, c, ns, W.out_hist[c0], nf);
                    }
                    else if (W.frame[c0 + 1] < cf || ns > W.score[c0 + 1]) ngf_enter(W, c0 + 1, ns, W.out_hist[c0], nf);
                }
            }
            else if (ns > wordthresh) ngs_save_bp(G, W, &entry, new_bp, new_s, cf, w, ns, W.out_hist[c0], 0);
        }
        for (int k = 0; k < ni; ++k) {
            const int c = c0 + 1 + k;
            if (W.frame[c] < cf) continue;
            if (W.best[c] > thresh) {
                const int ns = W.out_score[c] + pip;
                W.frame[c] = nf; W.word_active[w] = 1;
                if (ns > thresh) {
                    if (k == ni - 1) {
                        for (int r = rc0; r < c1; ++r) if (W.frame[r] < cf || ns > W.score[r]) ngf_enter(W, r, ns, W.out_hist[c], nf);
                    }
                    else if (W.frame[c + 1] < cf || ns > W.score[c + 1]) ngf_enter(W, c + 1, ns, W.out_hist[c], nf);
                }
            }
            else if (W.frame[c] != nf) ngf_clear_scores(G, W, c);
        }
        for (int c = rc0; c < c1; ++c) {
            if (W.frame[c] < cf) continue;
            if (W.best[c] > thresh) {
                W.frame[c] = nf; W.word_active[w] = 1;
                if (W.out_score[c] > wordthresh) ngs_save_bp(G, W, &entry, new_bp, new_s, cf, w, W.out_score[c], W.out_hist[c], c - rc0);
            }
            else if (W.frame[c] != nf) ngf_clear_scores(G, W, c);
        }
    }
    FSG_SYNC();
    const int bp0 = S->bpidx, bp1 = bp0 + n_new_bp;
    FSG_SYNC();
    FSG_IF_LEADER {
        S->bpidx = bp1; S->bss_head += n_new_bss;
        // best exit into silence (:745-752): first maximum over the frame's exits
        int best = FSG_WORST_SCORE, bb = 0;
        for (int b = bp0; b < bp1; ++b) {
            const fsg_wp e = W.bp + (size_t)b * NGS_BP_ROW;
            if (e[2] == G.finish_wid) continue;
            const int sc = e[9] == -1 ? e[4] : W.bss[e[5] + G.rs_cimap[((size_t)e[8] * G.n_ci + e[9]) * G.n_ci + G.sil]];
            if (sc > best) { best = sc; bb = b; }
        }
        S->silrc_score = best; S->silrc_bp = bb;
    }
    FSG_SYNC();
    // fwdflat_word_transition :643-782: targets = words with a node starting in [cf - win, cf + win)
    {
        int sf = cf - G.max_sf_win, ef = cf + G.max_sf_win;
        if (sf < 0) sf = 0;
        if (ef > T) ef = T;
        FSG_FOR(w, nwords) {
            if (W.wbase[w] < 0) continue;
            if (W.n_bp_in < 0) { if (!G.inlm[w]) continue; }
            else {                                                            // a surviving node starting in [sf, ef)?
                const int s0 = W.wstart[w], m = W.wn[w];
                int q = 0;
                while (q < m && W.nd_sf[s0 + q] < sf) ++q;
                if (!(q < m && W.nd_sf[s0 + q] < ef)) continue;
            }
            const int c0 = W.wbase[w], first = NGS_W(G, w, 0);
            const int ci2 = NGS_W(G, w, 3) ? G.sil : G.pron_ci[G.pron_off[w] + 1];
            for (int b = bp0; b < bp1; ++b) {
                const fsg_wp e = W.bp + (size_t)b * NGS_BP_ROW;
                if (e[2] == G.finish_wid) continue;
                int ns = e[9] == -1 ? e[4] : W.bss[e[5] + G.rs_cimap[((size_t)e[8] * G.n_ci + e[9]) * G.n_ci + first]];
                if (ns == FSG_WORST_SCORE) continue;
                ns = (int)FSG_FADD((float)ns, FSG_FMUL(G.lwf, (float)ngf_tg(G, NGS_W(G, w, 5), e[6], e[7])));
                ns += pip;
                if (!(ns > thresh)) continue;
                if (W.frame[c0] < cf || ns > W.score[c0]) {
                    ngf_enter(W, c0, ns, b, nf);
                    W.mss[c0] = G.ldiph[((size_t)first * G.n_ci + ci2) * G.n_ci + NGS_W(G, e[2], 1)];
                    W.word_active[w] = 1;
                }
            }
        }
    }
    FSG_SYNC();
    {
        const int ns_sil = S->silrc_score + G.silpen + pip, ns_fill = S->silrc_score + G.fillpen + pip;
        FSG_FOR(w, nwords) {
            if (w < G.filler_start || w > G.filler_end || !NGS_W(G, w, 3)) continue;
            const int ns = w == G.silence_wid ? ns_sil : ns_fill, c0 = W.wbase[w];
            if (!(ns > thresh && ns > FSG_WORST_SCORE)) continue;
            if (W.frame[c0] < cf || ns > W.score[c0]) { ngf_enter(W, c0, ns, S->silrc_bp, nf); W.word_active[w] = 1; }
        }
    }
    FSG_SYNC();
    FSG_FOR(j, nw) { const int c0 = W.wbase[awl[j]]; if (W.frame[c0] == cf) ngf_clear_scores(G, W, c0); }
    // next active word list :852-866
    const int nwd = S->nwd;
    FSG_FOR(k, nwd) { const int w = W.wordlist[k]; W.cnt[k] = (W.word_active[w] && w < G.start_wid) ? 1 : 0; }
    FSG_FOR(w, nwords) W.cnt3[w] = (w >= G.start_wid && W.word_active[w]) ? 1 : 0;
    FSG_SYNC();
    const int n1 = fsg_exscan(W.cnt, nwd, S->scan);
    const int n2 = fsg_exscan(W.cnt3, nwords, S->scan);
    FSG_FOR(k, nwd) { const int w = W.wordlist[k]; if (W.word_active[w] && w < G.start_wid) nawl[W.cnt[k]] = w; }
    FSG_FOR(w, nwords) if (w >= G.start_wid && W.word_active[w]) nawl[n1 + W.cnt3[w]] = w;
    FSG_SYNC();
    FSG_IF_LEADER { S->cur = nxt; S->n_awl = n1 + n2; S->n_done = cf + 1; }
    FSG_SYNC();
}
