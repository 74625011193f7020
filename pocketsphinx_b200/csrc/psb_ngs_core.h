// psb_ngs_core.h -- the first pass of n-gram decoding (ngram_search_fwdtree.c + the backpointer
// table half of ngram_search.c) for ONE utterance as block-wide data-parallel phases; same
// conventions as psb_fsg_core.h (FSG_FOR / FSG_SYNC / FSG_LEADER; compiled by nvcc into the kernel
// and by g++ into the test harness tests/emul/ngs_emul.cpp, which runs every loop to completion in
// ascending or descending "thread" order against the reference's golden backpointer tables).
//
// What the reference does sequentially, and the closed form used here:
//  * three coupled active sets -- static tree channels (roots by index, non-roots through an active
//    list), per-word right-context fan-out of word-final phones (allocated on demand: here a word
//    takes a block of channels from a per-utterance pool while any of them is allocated, "allocated"
//    being a flag per channel), permanent single-phone words.
//    All live in ONE channel index space: roots | non-roots | single-phone words | fan-out.
//  * prune_nonroot_chan (:800-878) walks the active list and, unlike the grammar search, its result
//    depends on the order: a child that failed the beam BEFORE its parent's turn has been cleared
//    and is entered unconditionally, one that comes later must be beaten on its state-0 score.
//    With pos = list position, par = the (single) parent:
//       enters(X)   = par survives, its exit score passes the phone beam (+ look-ahead penalty), and
//                     (X inactive  or  (pos X < pos par and X failed the beam)  or  score > X's)
//       X is appended to the next list by its parent if entered and not already there (it is there
//       iff it survived and stands earlier), by itself if it survives and was not entered earlier;
//       a failed X is cleared unless it was entered earlier; cleared-then-entered keeps only state 0.
//    Counts per position -> exclusive scan -> offsets, as in the grammar search.  An active child
//    applies its own entering (pull), an inactive one is written by its parent (push): one writer.
//  * last_phone_transition (:885-1035): every candidate word appears once per frame, so the
//    (end frame -> candidates) chains only organise a loop; each new candidate scans the backpointer
//    entries of its start frame itself (first maximum of exit score + trigram score).
//  * save_bp (ngram_search.c:378-497): one entry per (word, frame), created by the word's first exit
//    in right-context order, best score / path kept, every rc score on the side stack.  All exits
//    of a word are made by one thread; entry and stack offsets come from scans over the words.
//  * bptable_maxwpf (:1188-1238): "drop the worst until maxwpf remain" = rank by (score, index).
//  * word_transition (:1241-1430): per right-context phone the first maximum over the frame's
//    entries, then independent root / single-phone-word entries.
#pragma once
#include "psb_fsg_core.h"
#include "psb_lm_core.h"

#define NGS_BAD_SSID 0xffff
#define NGS_BP_ROW 10        /* frame, valid, wid, bp, score, s_idx, real_wid, prev_real_wid, last_phone, last2_phone */

struct NgsGraph {
    int n_words, n_root, n_nonroot, n_1ph, n_1ph_lm, n_ci, sil, n_lm, n_emit;
    int beam, pbeam, wbeam, lpbeam, lponlybeam, maxhmmpf, maxwpf, nwpen, pip, silpen, fillpen;
    int start_wid, finish_wid, silence_wid, filler_start, filler_end;
    int M, o_nonroot, o_1ph, o_rc, n_rcchan, LW;     // M: channels with per-utterance state; LW = scratch length (max of the list sizes) + 1
    int RB, n_blocks;          // fan-out pool: n_blocks blocks of RB channels (RB = widest fan-out); a word owns a block while any of
                               // its right-context channels is allocated.  State index o_rc + block * RB + rc, static index o_rc + wc_off[w] + rc
    const int32_t *roots;      // [n_root][5]   ciphone, ci2phone, penult_phn_wid, next, tmatid
    const int32_t *nonroot;    // [n_nonroot][6] ssid, tmatid, ciphone, penult_phn_wid, next, alt
    const int32_t *words;      // [n_words][8]  first, last, last2, single, filler, basewid, homophone, lmidx
    const int32_t *w1ph;       // [n_1ph]
    const int32_t *r1ph;       // [n_1ph][4]    ciphone, ci2phone, ssid, tmatid
    const int32_t *rs_n, *rs_ssid, *rs_cimap, *ldiph, *lm;
    const int32_t *wc_off;     // [n_words+1]   fan-out entries of every multi-phone word in the STATIC tables (tmatid, senid)
    const int32_t *w2h1;       // [n_words]     index of a single-phone word's permanent channel, or -1
    const int32_t *parent;     // [n_nonroot]   parent non-root id, or -(root id) - 1
    const int32_t *tmatid;     // [o_rc + n_rcchan]  by static index
    const int32_t *senid;      // [o_rc + n_rcchan][n_emit]   senones of the non-multiplexed channels (non-roots, fan-out)
    int use_lma;               // 1: trigram scores from the sorted-array LM below instead of the dense table `lm`
    LmArr lma;
};

struct NgsWork {
    fsg_wp score, hist, mss;                     // [n_emit][M]  (mss: per-state ssid of multiplexed channels)
    fsg_wp out_score, out_hist, best, frame;    // [M]
    fsg_wp alloc;                                  // [n_blocks * RB]
    fsg_wp wblock, free_stack;                     // [n_words] block of a word or -1; [n_blocks] free block ids
    fsg_wp pos, posf, eflag;                     // [n_nonroot]
    fsg_wp acl[2], awl[2];                        // [n_nonroot], [n_words]
    fsg_wp word_active, lt_sf, lt_dscr, lt_bp;  // [n_words]
    fsg_wp cand_wid, cand_score, cand_bp;        // [n_words]
    fsg_wp cnt, cnt2, cnt3, flag;               // [LW]
    fsg_wp brc_score, brc_path, brc_lc;          // [n_ci]
    fsg_wp bins;                                   // [256]
    fsg_wp bp, bss, bp_idx;                      // outputs: [bp_cap][10], [bss_cap], [T+1]
    const int32_t *pen;                              // [T][n_ci] the phone loop's penalties after each of ITS frames, or null
    int pl_window, T;                                // frame f of the search runs when the phone loop has seen frame min(f + window, T - 1)
    int bp_cap, bss_cap;
};

struct NgsScalars {
    fsg_int cur, n_acl, n_acl_nxt, n_awl, n_awl_nxt, n_cand;
    fsg_int best_all, best_last, best_score, last_phone_best, dynamic_beam, thresh, npth, lpth;
    fsg_int bpidx, bss_head, stop, error, n_done, k_nonfinish, ev_root, ev_last, n_free, renorm, norm;
    fsg_ll n_root_eval, n_nonroot_eval;
    int scan[34];
};

#define NGS_W(G, w, k) ((G).words[(size_t)(w) * 8 + (k)])

FSG_HD void ngs_work_carve(int32_t *b, const NgsGraph &G, NgsWork &W)
{
    const size_t M = (size_t)G.M, N = (size_t)G.n_emit, LW = (size_t)G.LW, nw = (size_t)G.n_words, nn = (size_t)G.n_nonroot + 1;
    W.score = b; b += N * M;  W.hist = b; b += N * M;  W.mss = b; b += N * M;
    W.out_score = b; b += M;  W.out_hist = b; b += M;  W.best = b; b += M;  W.frame = b; b += M;
    W.alloc = b; b += (size_t)G.n_blocks * G.RB + 1;  W.wblock = b; b += nw;  W.free_stack = b; b += (size_t)G.n_blocks + 1;
    W.pos = b; b += nn;  W.posf = b; b += nn;  W.eflag = b; b += nn;
    W.acl[0] = b; b += nn;  W.acl[1] = b; b += nn;  W.awl[0] = b; b += nw + 1;  W.awl[1] = b; b += nw + 1;
    W.word_active = b; b += nw;  W.lt_sf = b; b += nw;  W.lt_dscr = b; b += nw;  W.lt_bp = b; b += nw;
    W.cand_wid = b; b += nw + 1;  W.cand_score = b; b += nw + 1;  W.cand_bp = b; b += nw + 1;
    W.cnt = b; b += LW;  W.cnt2 = b; b += LW;  W.cnt3 = b; b += LW;  W.flag = b; b += LW;
    W.brc_score = b; b += G.n_ci;  W.brc_path = b; b += G.n_ci;  W.brc_lc = b; b += G.n_ci;
    W.bins = b;
}

FSG_HDH size_t ngs_work_words(const NgsGraph &G)
{
    const size_t M = (size_t)G.M, N = (size_t)G.n_emit, LW = (size_t)G.LW, nw = (size_t)G.n_words, nn = (size_t)G.n_nonroot + 1;
    return 3 * N * M + 4 * M + (size_t)G.n_blocks * G.RB + 1 + nw + (size_t)G.n_blocks + 1 + 5 * nn + 2 * (nw + 1) + 4 * nw + 3 * (nw + 1) + 4 * LW + 3 * (size_t)G.n_ci + 256;
}

FSG_HD int ngs_nrc(const NgsGraph &G, int w) { return G.wc_off[w + 1] - G.wc_off[w]; }
FSG_HD int ngs_pl(const NgsWork &W, const NgsGraph &G, int f, int ci)          /* phone_loop_search_score at search frame f */
{
    if (!W.pen) return 0;
    int t = f + W.pl_window;                                                  /* ps_search_forward / ps_end_utt, pocketsphinx.c:1172-1195, 1329-1333 */
    if (t > W.T - 1) t = W.T - 1;
    return W.pen[(size_t)t * G.n_ci + ci];
}

FSG_HD void ngs_normalize(const NgsGraph &G, const NgsWork &W, int c, int norm)  /* hmm_normalize, hmm.c:206-217 */
{
    for (int s = 0; s < G.n_emit; ++s) {
        const int v = W.score[s * G.M + c];
        if (v > FSG_WORST_SCORE) W.score[s * G.M + c] = v - norm;
    }
    const int o = W.out_score[c];
    if (o > FSG_WORST_SCORE) W.out_score[c] = o - norm;
}

FSG_HD void ngs_clear(const NgsGraph &G, const NgsWork &W, int c)               /* hmm_clear */
{
    for (int s = 0; s < G.n_emit; ++s) { W.score[s * G.M + c] = FSG_WORST_SCORE; W.hist[s * G.M + c] = -1; }
    W.out_score[c] = FSG_WORST_SCORE; W.out_hist[c] = -1; W.best[c] = FSG_WORST_SCORE; W.frame[c] = -1;
}

FSG_HD void ngs_enter(const NgsWork &W, int c, int score, int hist, int nf)     /* hmm_enter */
{
    W.score[c] = score; W.hist[c] = hist; W.frame[c] = nf;
}

FSG_HDH int ngs_tg(const NgsGraph &G, int w, int h1, int h2)
{
    if (G.use_lma) return lm_tg_score(G.lma, w, h1, h2) >> 10;           /* >> SENSCR_SHIFT */
    const int n = G.n_lm + 1;
    const int a = NGS_W(G, w, 7), b = h1 < 0 ? 0 : NGS_W(G, h1, 7) + 1, c = h2 < 0 ? 0 : NGS_W(G, h2, 7) + 1;
    return G.lm[((size_t)a * n + b) * n + c];
}

FSG_HD int ngs_exit_score(const NgsGraph &G, const NgsWork &W, const fsg_wp e, int rcphone)   /* ngram_search.c:655-676 */
{
    if (e[9] == -1) return e[4];
    return W.bss[e[5] + G.rs_cimap[((size_t)e[8] * G.n_ci + e[9]) * G.n_ci + rcphone]];
}

template <class GraphT, class WorkT>
FSG_HD void ngs_set_real_wid(const GraphT &G, const WorkT &W, int bp)                       /* :343-373 */
{
    fsg_wp e = W.bp + (size_t)bp * NGS_BP_ROW;
    const fsg_wp prev = e[3] == -1 ? fsg_wp(nullptr) : W.bp + (size_t)e[3] * NGS_BP_ROW;
    if (NGS_W(G, e[2], 4)) {
        if (prev) { e[6] = prev[6]; e[7] = prev[7]; }
        else { e[6] = NGS_W(G, e[2], 5); e[7] = -1; }
    }
    else {
        e[6] = NGS_W(G, e[2], 5);
        e[7] = prev ? prev[6] : -1;
    }
}

// One word's exits of one frame, made in order by one thread (save_bp).  Shared with the second pass
// (psb_ngf_core.h): any graph / work pair with words, rs_n, n_ci and bp, bss.  *entry = -1 before the
// first; new entries take index new_bp / stack offset new_s (from the scans).
template <class GraphT, class WorkT>
FSG_HD void ngs_save_bp(const GraphT &G, const WorkT &W, int *entry, int new_bp, int new_s, int frame, int w,
                        int score, int path, int rc)
{
    if (*entry != -1) {
        fsg_wp e = W.bp + (size_t)*entry * NGS_BP_ROW;
        if (e[4] < score) {
            if (e[3] != path) {
                // The reference re-derives the entry's LM state here BEFORE it moves to the new path
                // (:420-440): it lags one update behind, which is visible after a second change.
                const fsg_wp po = e[3] == -1 ? fsg_wp(nullptr) : W.bp + (size_t)e[3] * NGS_BP_ROW;
                const fsg_wp pn = path == -1 ? fsg_wp(nullptr) : W.bp + (size_t)path * NGS_BP_ROW;
                const int a0 = po ? po[7] : -1, a1 = po ? po[6] : -1, b0 = pn ? pn[7] : -1, b1 = pn ? pn[6] : -1;
                if (a0 != b0 || a1 != b1) ngs_set_real_wid(G, W, *entry);
                e[3] = path;
            }
            e[4] = score;
        }
        if (e[5] != -1) W.bss[e[5] + rc] = score;
    }
    else {
        fsg_wp e = W.bp + (size_t)new_bp * NGS_BP_ROW;
        int rcsize;
        *entry = new_bp;
        e[2] = w; e[0] = frame; e[3] = path; e[4] = score; e[5] = new_s; e[1] = 1;
        e[8] = NGS_W(G, w, 1);
        if (NGS_W(G, w, 3)) { e[9] = -1; e[5] = -1; rcsize = 0; }
        else { e[9] = NGS_W(G, w, 2); rcsize = G.rs_n[(size_t)e[8] * G.n_ci + e[9]]; }
        for (int i = 0; i < rcsize; ++i) W.bss[new_s + i] = FSG_WORST_SCORE;
        if (rcsize) W.bss[new_s + rc] = score;
        ngs_set_real_wid(G, W, new_bp);
    }
}

FSG_HD void ngs_start(const NgsGraph &G, const NgsWork &W, NgsScalars *S)        /* ngram_fwdtree_start :470-520 */
{
    FSG_FOR(c, G.M) {
        ngs_clear(G, W, c);
        for (int s = 0; s < G.n_emit; ++s) W.mss[s * G.M + c] = s == 0 ? 0 : NGS_BAD_SSID;
        if (c >= G.o_1ph && c < G.o_rc) W.mss[c] = G.r1ph[(c - G.o_1ph) * 4 + 2];
    }
    FSG_FOR(i, G.n_blocks * G.RB) W.alloc[i] = 0;
    FSG_FOR(i, G.n_blocks) W.free_stack[i] = i;
    FSG_FOR(w, G.n_words) { W.lt_sf[w] = -1; W.lt_dscr[w] = 0; W.lt_bp[w] = 0; W.word_active[w] = 0; W.wblock[w] = -1; }
    FSG_FOR(i, G.n_nonroot) { W.pos[i] = -1; W.posf[i] = -2; W.eflag[i] = 0; }
    FSG_IF_LEADER {
        S->cur = 0; S->n_acl = S->n_acl_nxt = S->n_awl = S->n_awl_nxt = S->n_cand = 0;
        S->best_score = 0; S->last_phone_best = 0; S->dynamic_beam = G.beam;
        S->bpidx = 0; S->bss_head = 0; S->stop = 0; S->error = 0; S->n_done = 0;
        S->n_root_eval = 0; S->n_nonroot_eval = 0; S->n_free = G.n_blocks;
    }
    FSG_SYNC();
    FSG_IF_LEADER ngs_enter(W, G.o_1ph + G.w2h1[G.start_wid], 0, -1, 0);
    FSG_SYNC();
}

// does non-root X's parent enter it this frame?  Evaluated on the state evaluation left (phase E1).
FSG_HD bool ngs_parent_enters(const NgsGraph &G, const NgsWork &W, const NgsScalars *S, int X, int f, bool x_active, int kX, bool x_surv)
{
    const int par = G.parent[X];
    if (par < 0) return false;                                   // root parents act in phase D
    if (W.posf[par] != f) return false;
    const int cp = G.o_nonroot + par;
    if (!(W.best[cp] > S->thresh)) return false;
    const int nps = W.out_score[cp] + G.pip;
    if (!(nps + ngs_pl(W, G, f, G.nonroot[X * 6 + 2]) > S->npth)) return false;
    if (!x_active) return true;
    if (kX < W.pos[par] && !x_surv) return true;
    return nps > W.score[G.o_nonroot + X];
}

template <class Eval>
FSG_HD void ngs_step(const NgsGraph &G, const NgsWork &W, NgsScalars *S, int f, Eval &eval)
{
    const int nf = f + 1, cur = S->cur, nxt = cur ^ 1, n_acl = S->n_acl, n_awl = S->n_awl, M = G.M;
    fsg_wp acl = W.acl[cur], nacl = W.acl[nxt], awl = W.awl[cur], nawl = W.awl[nxt];
    // ---- ngram_fwdtree_search :1454-1482
    FSG_IF_LEADER {
        W.bp_idx[f] = S->bpidx;
        if (S->best_score <= FSG_WORST_SCORE) S->stop = 1;
        S->renorm = 0;
        if (!S->stop && S->best_score + 2 * G.beam < FSG_WORST_SCORE) { S->renorm = 1; S->norm = S->best_score; }
        S->best_all = FSG_WORST_SCORE; S->best_last = FSG_WORST_SCORE; S->ev_root = 0; S->ev_last = 0;
    }
    FSG_SYNC();
    if (S->stop || S->error) return;
    if (S->renorm) {
        // ---- renormalize_scores :566-602: hmm_normalize on every channel that is about to be evaluated
        // (one writer per channel; nothing else is read in this phase)
        const int norm = S->norm;
        FSG_FOR(i, G.n_root) if (W.frame[i] == f) ngs_normalize(G, W, i, norm);
        FSG_FOR(k, n_acl) ngs_normalize(G, W, G.o_nonroot + acl[k], norm);
        FSG_FOR(j, n_awl) {
            const int w = awl[j], blk = W.wblock[w] * G.RB, nrc = ngs_nrc(G, w);
            for (int r = 0; r < nrc; ++r)
                if (W.alloc[blk + r]) ngs_normalize(G, W, G.o_rc + blk + r, norm);
        }
        FSG_FOR(i, G.n_1ph) if (W.frame[G.o_1ph + i] == f) ngs_normalize(G, W, G.o_1ph + i, norm);
        FSG_SYNC();
    }
    // ---- evaluate_channels :702-716
    FSG_FOR(i, G.n_root) if (W.frame[i] == f) { FSG_ATOMIC_MAX(&S->best_all, eval(W, i, true, i)); FSG_ATOMIC_ADD(&S->ev_root, 1); }
    FSG_FOR(k, n_acl) {
        const int id = acl[k];
        W.pos[id] = k; W.posf[id] = f;
        FSG_ATOMIC_MAX(&S->best_all, eval(W, G.o_nonroot + id, false, G.o_nonroot + id));
    }
    FSG_FOR(j, n_awl) {
        const int w = awl[j];
        int k = 0;
        W.word_active[w] = 0;
        const int blk = W.wblock[w] * G.RB, nrc = ngs_nrc(G, w);
        for (int r = 0; r < nrc; ++r)
            if (W.alloc[blk + r]) { FSG_ATOMIC_MAX(&S->best_last, eval(W, G.o_rc + blk + r, false, G.o_rc + G.wc_off[w] + r)); ++k; }
        FSG_ATOMIC_ADD(&S->ev_last, k);
    }
    FSG_FOR(i, G.n_1ph) {
        const int c = G.o_1ph + i;
        if (W.frame[c] < f) continue;
        const int sc = eval(W, c, true, c);
        if (G.w1ph[i] != G.finish_wid) FSG_ATOMIC_MAX(&S->best_last, sc);
        FSG_ATOMIC_ADD(&S->ev_last, 1);
    }
    FSG_SYNC();
    // ---- prune_channels :1130-1180: beams
    FSG_IF_LEADER {
        S->best_score = S->best_all > S->best_last ? S->best_all : S->best_last;
        S->last_phone_best = S->best_last;
        S->n_root_eval += S->ev_root; S->n_nonroot_eval += n_acl + S->ev_last;
        S->dynamic_beam = G.beam;
        S->n_cand = 0;
    }
    FSG_SYNC();
    if (G.maxhmmpf != -1 && S->n_root_eval + S->n_nonroot_eval > G.maxhmmpf) {
        const int bw = -G.beam / 256;
        FSG_FOR(b, 256) W.bins[b] = 0;
        FSG_SYNC();
        FSG_FOR(i, G.n_root) { int b = (S->best_score - W.best[i]) / bw; if (b >= 256) b = 255; FSG_ATOMIC_ADD_AT(W.bins, b, 1); }
        FSG_FOR(k, n_acl) { int b = (S->best_score - W.best[G.o_nonroot + acl[k]]) / bw; if (b >= 256) b = 255; FSG_ATOMIC_ADD_AT(W.bins, b, 1); }
        FSG_SYNC();
        FSG_IF_LEADER {
            int i, nh = 0;
            for (i = 0; i < 256; ++i) { nh += W.bins[i]; if (nh > G.maxhmmpf) break; }
            S->dynamic_beam = -(i * bw);
        }
        FSG_SYNC();
    }
    FSG_IF_LEADER { S->thresh = S->best_score + S->dynamic_beam; S->npth = S->best_score + G.pbeam; S->lpth = S->best_score + G.lpbeam; }
    FSG_SYNC();
    const int thresh = S->thresh, npth = S->npth, lpth = S->lpth;
    // ---- prune_root_chan :723-794.  cnt = children entered, cnt2 = last-phone candidates
    FSG_FOR(i, G.n_root) {
        int c1 = 0, c2 = 0, fl = 0;
        if (W.frame[i] >= f && W.best[i] > thresh) {
            const int nps = W.out_score[i] + G.pip;
            fl = 1;
            if (W.pen || nps > npth)
                for (int c = G.roots[i * 5 + 3]; c >= 0; c = G.nonroot[c * 6 + 5])
                    if (nps + ngs_pl(W, G, f, G.nonroot[c * 6 + 2]) > npth && (W.frame[G.o_nonroot + c] < f || nps > W.score[G.o_nonroot + c])) ++c1;
            if (W.pen || nps > lpth)
                for (int w = G.roots[i * 5 + 2]; w >= 0; w = NGS_W(G, w, 6))
                    if (nps + ngs_pl(W, G, f, NGS_W(G, w, 1)) > lpth) ++c2;
        }
        W.cnt[i] = c1; W.cnt2[i] = c2; W.flag[i] = fl;
    }
    FSG_FOR(i, G.n_nonroot) W.eflag[i] = 0;
    FSG_SYNC();
    const int n_from_roots = fsg_exscan(W.cnt, G.n_root, S->scan);
    const int n_cand_roots = fsg_exscan(W.cnt2, G.n_root, S->scan);
    FSG_FOR(i, G.n_root) {
        if (!W.flag[i]) continue;
        const int nps = W.out_score[i] + G.pip;
        int o1 = W.cnt[i], o2 = W.cnt2[i];
        W.frame[i] = nf;
        if (W.pen || nps > npth)
            for (int c = G.roots[i * 5 + 3]; c >= 0; c = G.nonroot[c * 6 + 5]) {
                const int cc = G.o_nonroot + c;
                if (nps + ngs_pl(W, G, f, G.nonroot[c * 6 + 2]) > npth && (W.frame[cc] < f || nps > W.score[cc])) {
                    ngs_enter(W, cc, nps, W.out_hist[i], nf);
                    W.eflag[c] = 1;
                    nacl[o1++] = c;
                }
            }
        if (W.pen || nps > lpth)
            for (int w = G.roots[i * 5 + 2]; w >= 0; w = NGS_W(G, w, 6))
                if (nps + ngs_pl(W, G, f, NGS_W(G, w, 1)) > lpth) {
                    W.cand_wid[o2] = w; W.cand_score[o2] = nps - G.nwpen; W.cand_bp[o2] = W.out_hist[i]; ++o2;
                }
    }
    FSG_SYNC();
    // ---- prune_nonroot_chan :800-878.  E1: every active node decides whether its (non-root) parent enters it
    FSG_FOR(k, n_acl) {
        const int X = acl[k];
        if (G.parent[X] >= 0)
            W.eflag[X] = ngs_parent_enters(G, W, S, X, f, true, k, W.best[G.o_nonroot + X] > thresh) ? 1 : 0;
    }
    FSG_SYNC();
    // E2: counts.  flag bits: 1 self-append, 2 transitions to children, 4 candidates
    FSG_FOR(k, n_acl) {
        const int X = acl[k], cx = G.o_nonroot + X;
        int c1 = 0, c2 = 0, fl = 0;
        if (W.best[cx] > thresh) {
            const int par = G.parent[X];
            const bool earlier = W.eflag[X] && (par < 0 || W.pos[par] < k);
            const int nps = W.out_score[cx] + G.pip;
            if (!earlier) { c1 = 1; fl |= 1; }
            if (W.pen || nps > npth) {
                fl |= 2;
                for (int c = G.nonroot[X * 6 + 4]; c >= 0; c = G.nonroot[c * 6 + 5]) {
                    const bool c_act = W.posf[c] == f;
                    const bool ent = c_act ? W.eflag[c] != 0 : ngs_parent_enters(G, W, S, c, f, false, 0, false);
                    if (ent && !(c_act && W.pos[c] < k && W.best[G.o_nonroot + c] > thresh)) ++c1;
                }
            }
            if (W.pen || nps > lpth) {
                fl |= 4;
                for (int w = G.nonroot[X * 6 + 3]; w >= 0; w = NGS_W(G, w, 6))
                    if (nps + ngs_pl(W, G, f, NGS_W(G, w, 1)) > lpth) ++c2;
            }
        }
        W.cnt[k] = c1; W.cnt2[k] = c2; W.flag[k] = fl;
    }
    FSG_SYNC();
    const int n_from_nonroot = fsg_exscan(W.cnt, n_acl, S->scan);
    const int n_cand_nonroot = fsg_exscan(W.cnt2, n_acl, S->scan);
    // E3: apply.  Active nodes write themselves (clear, then their own entering); parents write inactive children.
    FSG_FOR(k, n_acl) {
        const int X = acl[k], cx = G.o_nonroot + X, fl = W.flag[k], par = G.parent[X];
        const bool surv = W.best[cx] > thresh;
        const bool earlier = W.eflag[X] && (par < 0 || W.pos[par] < k);
        const int my_out = W.out_score[cx], my_hist = W.out_hist[cx];
        int o1 = n_from_roots + W.cnt[k], o2 = n_cand_roots + W.cnt2[k];
        if (fl & 1) nacl[o1++] = X;
        if (fl & 2) {
            const int nps = my_out + G.pip;
            for (int c = G.nonroot[X * 6 + 4]; c >= 0; c = G.nonroot[c * 6 + 5]) {
                const bool c_act = W.posf[c] == f;
                const bool ent = c_act ? W.eflag[c] != 0 : ngs_parent_enters(G, W, S, c, f, false, 0, false);
                if (!ent) continue;
                if (!(c_act && W.pos[c] < k && W.best[G.o_nonroot + c] > thresh)) nacl[o1++] = c;
                if (!c_act) ngs_enter(W, G.o_nonroot + c, nps, my_hist, nf);
            }
        }
        if (fl & 4) {
            const int nps = my_out + G.pip;
            for (int w = G.nonroot[X * 6 + 3]; w >= 0; w = NGS_W(G, w, 6))
                if (nps + ngs_pl(W, G, f, NGS_W(G, w, 1)) > lpth) {
                    W.cand_wid[o2] = w; W.cand_score[o2] = nps - G.nwpen; W.cand_bp[o2] = my_hist; ++o2;
                }
        }
        if (surv) W.frame[cx] = nf;
        else if (!earlier) ngs_clear(G, W, cx);
        if (W.eflag[X] && par >= 0) {                                         // entered by a non-root parent: pull
            const int cp = G.o_nonroot + par;
            ngs_enter(W, cx, W.out_score[cp] + G.pip, W.out_hist[cp], nf);
        }
    }
    FSG_SYNC();
    const int n_cand = n_cand_roots + n_cand_nonroot;
    // ---- last_phone_transition :885-1035
    FSG_IF_LEADER { S->n_acl_nxt = n_from_roots + n_from_nonroot; S->n_cand = n_cand; S->best_all = S->last_phone_best; }
    FSG_FOR(i, n_cand) {
        const int w = W.cand_wid[i], bp0 = W.cand_bp[i];
        if (bp0 == -1) continue;
        const fsg_wp e0 = W.bp + (size_t)bp0 * NGS_BP_ROW;
        const int ef = e0[0], first = NGS_W(G, w, 0);
        W.cand_score[i] -= ngs_exit_score(G, W, e0, first);
        if (W.lt_sf[w] != ef + 1) {
            int dbest = FSG_WORST_SCORE, bbest = W.lt_bp[w];
            for (int bp = W.bp_idx[ef]; bp < W.bp_idx[ef + 1]; ++bp) {
                const fsg_wp e = W.bp + (size_t)bp * NGS_BP_ROW;
                if (!e[1]) continue;
                int dscr = ngs_exit_score(G, W, e, first);
                if (dscr > FSG_WORST_SCORE) dscr += ngs_tg(G, NGS_W(G, w, 5), e[6], e[7]);
                if (dscr > dbest) { dbest = dscr; bbest = bp; }
            }
            W.lt_dscr[w] = dbest; W.lt_bp[w] = bbest; W.lt_sf[w] = ef + 1;
        }
    }
    FSG_SYNC();
    FSG_FOR(i, n_cand) {
        const int w = W.cand_wid[i];
        W.cand_score[i] += W.lt_dscr[w];
        W.cand_bp[i] = W.lt_bp[w];
        FSG_ATOMIC_MAX(&S->best_all, W.cand_score[i]);
    }
    FSG_SYNC();
    FSG_IF_LEADER S->last_phone_best = S->best_all;
    FSG_SYNC();
    const int th_lp = S->last_phone_best + G.lponlybeam;
    // words entering their last phone without a fan-out block take one from the pool (order is immaterial)
    FSG_FOR(i, n_cand) W.cnt2[i] = (W.cand_score[i] > th_lp && W.wblock[W.cand_wid[i]] < 0) ? 1 : 0;
    FSG_SYNC();
    const int n_need = fsg_exscan(W.cnt2, n_cand, S->scan);
    if (n_need > S->n_free) {
        FSG_IF_LEADER S->error = 3;                                           // fan-out pool exhausted
        FSG_SYNC();
        return;
    }
    FSG_FOR(i, n_cand) {
        const int w = W.cand_wid[i];
        if (W.cand_score[i] > th_lp && W.wblock[w] < 0) W.wblock[w] = W.free_stack[S->n_free - 1 - W.cnt2[i]];
    }
    FSG_SYNC();
    FSG_IF_LEADER S->n_free -= n_need;
    FSG_FOR(i, n_cand) {
        int k = 0;
        if (W.cand_score[i] > th_lp) {
            const int w = W.cand_wid[i], sc = W.cand_score[i], blk = W.wblock[w] * G.RB, nrc = ngs_nrc(G, w);
            for (int r = 0; r < nrc; ++r) {
                const int c = G.o_rc + blk + r;
                if (!W.alloc[blk + r]) { ngs_clear(G, W, c); W.alloc[blk + r] = 1; }             // ngram_search_alloc_all_rc
                if (W.frame[c] < f || sc > W.score[c]) { ngs_enter(W, c, sc, W.cand_bp[i], nf); ++k; }
            }
            if (k > 0) W.word_active[w] = 1;
        }
        W.cnt[i] = k > 0 ? 1 : 0;
    }
    FSG_SYNC();
    const int n_awl_lp = fsg_exscan(W.cnt, n_cand, S->scan);
    FSG_FOR(i, n_cand) {
        const bool mine = (i + 1 < n_cand ? W.cnt[i + 1] : n_awl_lp) > W.cnt[i];
        if (mine) nawl[W.cnt[i]] = W.cand_wid[i];
    }
    FSG_SYNC();
    // ---- prune_word_chan :1042-1126: items = active words, then the single-phone words
    const int newword_thresh = S->last_phone_best + G.wbeam, lp_thresh = S->last_phone_best + G.lponlybeam;
    const int n_items = n_awl + G.n_1ph;
    FSG_FOR(j, n_items) {
        int has_exit = 0, rcsize = 0, k = 0;
        if (j < n_awl) {
            const int w = awl[j], blk = W.wblock[w] * G.RB, nrc = ngs_nrc(G, w);
            for (int r = 0; r < nrc; ++r) {
                const int c = G.o_rc + blk + r;
                if (!W.alloc[blk + r]) continue;
                if (W.best[c] > lp_thresh) { ++k; if (W.out_score[c] > newword_thresh) has_exit = 1; }
            }
            if (has_exit) rcsize = ngs_nrc(G, w);
            W.cnt3[j] = (k > 0 && !W.word_active[w]) ? 1 : 0;
        }
        else {
            const int c = G.o_1ph + (j - n_awl);
            if (W.frame[c] >= f && W.best[c] > lp_thresh && W.out_score[c] > newword_thresh) has_exit = 1;
            W.cnt3[j] = 0;
        }
        W.cnt[j] = has_exit; W.cnt2[j] = rcsize;
    }
    FSG_SYNC();
    const int n_new_bp = fsg_exscan(W.cnt, n_items, S->scan);
    const int n_new_bss = fsg_exscan(W.cnt2, n_items, S->scan);
    const int n_awl_wc = fsg_exscan(W.cnt3, n_items, S->scan);
    if (S->bpidx + n_new_bp > W.bp_cap || S->bss_head + n_new_bss > W.bss_cap) {
        FSG_IF_LEADER S->error = 1;
        FSG_SYNC();
        return;
    }
    FSG_FOR(j, n_items) {
        int entry = -1;
        const int new_bp = S->bpidx + W.cnt[j], new_s = S->bss_head + W.cnt2[j];
        if (j < n_awl) {
            const int w = awl[j], blk = W.wblock[w] * G.RB, nrc = ngs_nrc(G, w);
            int k = 0, left = 0;
            for (int r = 0; r < nrc; ++r) {
                const int c = G.o_rc + blk + r;
                if (!W.alloc[blk + r]) continue;
                if (W.best[c] > lp_thresh) {
                    W.frame[c] = nf; ++k;
                    if (W.out_score[c] > newword_thresh) ngs_save_bp(G, W, &entry, new_bp, new_s, f, w, W.out_score[c], W.out_hist[c], r);
                }
                else if (W.frame[c] != nf) W.alloc[blk + r] = 0;
                left += W.alloc[blk + r];
            }
            if (k > 0 && !W.word_active[w]) { nawl[n_awl_lp + W.cnt3[j]] = w; W.word_active[w] = 1; }
            W.flag[j] = left == 0 ? 1 : 0;                                   // every channel freed: the block goes back
        }
        else {
            const int i = j - n_awl, c = G.o_1ph + i;
            if (W.frame[c] >= f && W.best[c] > lp_thresh) {
                W.frame[c] = nf;
                if (W.out_score[c] > newword_thresh) ngs_save_bp(G, W, &entry, new_bp, new_s, f, G.w1ph[i], W.out_score[c], W.out_hist[c], 0);
            }
        }
    }
    FSG_SYNC();
    // words all of whose fan-out channels were freed return their block to the pool
    FSG_FOR(j, n_awl) W.cnt3[j] = W.flag[j];
    FSG_SYNC();
    const int n_rel = fsg_exscan(W.cnt3, n_awl, S->scan);
    FSG_FOR(j, n_awl) {
        if (!W.flag[j]) continue;
        const int w = awl[j];
        W.free_stack[S->n_free + W.cnt3[j]] = W.wblock[w];
        W.wblock[w] = -1;
    }
    FSG_SYNC();
    FSG_IF_LEADER S->n_free += n_rel;
    const int bp0 = S->bpidx, bp1 = bp0 + n_new_bp;
#ifndef NGS_TEST_INJECT_RACE                                                  /* tests/test_emul_racecheck.py removes this barrier to prove the detector sees it */
    FSG_SYNC();                                                               // everyone has read bpidx before it moves
#endif
    FSG_IF_LEADER { S->bpidx = bp1; S->bss_head += n_new_bss; S->n_awl_nxt = n_awl_lp + n_awl_wc; S->k_nonfinish = 0; }
    FSG_SYNC();
    // ---- bptable_maxwpf :1188-1238
    if (G.maxwpf != -1 && G.maxwpf != G.n_words) {
        FSG_IF_LEADER {                                                     // fillers: only the best stays valid
            int n = 0, bestscr = INT_MIN, bestbp = -1;
            for (int bp = bp0; bp < bp1; ++bp) {
                fsg_wp e = W.bp + (size_t)bp * NGS_BP_ROW;
                if (NGS_W(G, e[2], 4)) { if (e[4] > bestscr) { bestscr = e[4]; bestbp = bp; } e[1] = 0; ++n; }
            }
            if (bestbp >= 0) { W.bp[(size_t)bestbp * NGS_BP_ROW + 1] = 1; --n; }
            S->scan[32] = (bp1 - bp0) - n - G.maxwpf;                           // how many of the worst to drop
        }
        FSG_SYNC();
        const int n_drop = S->scan[32];
        if (n_drop > 0) {
            FSG_FOR(x, bp1 - bp0) {
                const fsg_wp e = W.bp + (size_t)(bp0 + x) * NGS_BP_ROW;
                int rank = -1;
                if (e[1]) {
                    rank = 0;
                    for (int y = 0; y < bp1 - bp0; ++y) {
                        const fsg_wp o = W.bp + (size_t)(bp0 + y) * NGS_BP_ROW;
                        if (y != x && o[1] && (o[4] < e[4] || (o[4] == e[4] && y < x))) ++rank;
                    }
                }
                W.cnt[x] = (rank >= 0 && rank < n_drop) ? 1 : 0;
            }
            FSG_SYNC();
            FSG_FOR(x, bp1 - bp0) if (W.cnt[x]) W.bp[(size_t)(bp0 + x) * NGS_BP_ROW + 1] = 0;
            FSG_SYNC();
        }
    }
    // ---- word_transition :1241-1430
    FSG_FOR(rc, G.n_ci) {
        int best = FSG_WORST_SCORE, path = W.brc_path[rc], lc = W.brc_lc[rc], k = 0;
        for (int bp = bp0; bp < bp1; ++bp) {
            const fsg_wp e = W.bp + (size_t)bp * NGS_BP_ROW;
            if (e[2] == G.finish_wid) continue;
            ++k;
            const int sc = e[9] == -1 ? e[4] : W.bss[e[5] + G.rs_cimap[((size_t)e[8] * G.n_ci + e[9]) * G.n_ci + rc]];
            if (sc > best) { best = sc; path = bp; lc = e[8]; }
        }
        W.brc_score[rc] = best; W.brc_path[rc] = path; W.brc_lc[rc] = lc;
        if (rc == 0 && k > 0) S->k_nonfinish = k;
    }
    FSG_SYNC();
    if (S->k_nonfinish > 0) {
        const int th = S->best_score + S->dynamic_beam, nc = G.n_ci;
        FSG_FOR(i, G.n_root) {
            const int ci = G.roots[i * 5], ci2 = G.roots[i * 5 + 1];
            const int ns = W.brc_score[ci] + G.nwpen + G.pip;
            if (ns + ngs_pl(W, G, f, ci) > th && (W.frame[i] < f || ns > W.score[i])) {
                ngs_enter(W, i, ns, W.brc_path[ci], nf);
                W.mss[i] = G.ldiph[((size_t)ci * nc + ci2) * nc + W.brc_lc[ci]];
            }
        }
        FSG_FOR(i, G.n_1ph) {
            const int w = G.w1ph[i], c = G.o_1ph + i, ci = G.r1ph[i * 4];
            if (i < G.n_1ph_lm && w != G.start_wid) {
                int dbest = INT_MIN, bbest = W.lt_bp[w];
                for (int bp = bp0; bp < bp1; ++bp) {
                    const fsg_wp e = W.bp + (size_t)bp * NGS_BP_ROW;
                    if (!e[1]) continue;
                    int ns = ngs_exit_score(G, W, e, NGS_W(G, w, 0));
                    if (ns != FSG_WORST_SCORE) ns += ngs_tg(G, NGS_W(G, w, 5), e[6], e[7]);
                    if (ns > dbest) { dbest = ns; bbest = bp; }
                }
                W.lt_dscr[w] = dbest; W.lt_bp[w] = bbest;
                const int ns = (int)((unsigned)dbest + (unsigned)G.pip);
                if ((int)((unsigned)ns + (unsigned)ngs_pl(W, G, f, ci)) > th && (W.frame[c] < f || ns > W.score[c])) {
                    ngs_enter(W, c, ns, bbest, nf);
                    W.mss[c] = G.ldiph[((size_t)ci * nc + G.r1ph[i * 4 + 1]) * nc + NGS_W(G, W.bp[(size_t)bbest * NGS_BP_ROW + 2], 1)];
                }
            }
            else if (i < G.n_1ph_lm) {                                           // <s>: only the cache is refreshed (:1339-1362)
                int dbest = INT_MIN, bbest = W.lt_bp[w];
                for (int bp = bp0; bp < bp1; ++bp) {
                    const fsg_wp e = W.bp + (size_t)bp * NGS_BP_ROW;
                    if (!e[1]) continue;
                    int ns = ngs_exit_score(G, W, e, NGS_W(G, w, 0));
                    if (ns != FSG_WORST_SCORE) ns += ngs_tg(G, NGS_W(G, w, 5), e[6], e[7]);
                    if (ns > dbest) { dbest = ns; bbest = bp; }
                }
                W.lt_dscr[w] = dbest; W.lt_bp[w] = bbest;
            }
        }
        FSG_SYNC();
        // <sil> and the other fillers (:1386-1428); they come after the LM words and may overwrite them (</s>)
        FSG_FOR(w, G.n_words) {
            if (w < G.filler_start || w > G.filler_end || G.w2h1[w] < 0 || w == G.start_wid) continue;
            const int c = G.o_1ph + G.w2h1[w];
            const int ns = W.brc_score[G.sil] + (w == G.silence_wid ? G.silpen : G.fillpen) + G.pip;
            if (ns + ngs_pl(W, G, f, G.r1ph[G.w2h1[w] * 4]) > th && (W.frame[c] < f || ns > W.score[c]))
                ngs_enter(W, c, ns, W.brc_path[G.sil], nf);
        }
        FSG_SYNC();
    }
    // ---- deactivate_channels :1432-1451
    FSG_FOR(i, G.n_root) if (W.frame[i] == f) ngs_clear(G, W, i);
    FSG_FOR(i, G.n_1ph) if (W.frame[G.o_1ph + i] == f) ngs_clear(G, W, G.o_1ph + i);
    FSG_IF_LEADER { S->cur = nxt; S->n_acl = S->n_acl_nxt; S->n_awl = S->n_awl_nxt; S->n_done = f + 1; }
    FSG_SYNC();
    (void)M;
}
