// psb_fsg_host.h -- host-side preparation of the grammar search: checks the flattened lextree a
// caller hands to psb_fsg_batch_device and lays it out as one int32 block for the device
// (FsgGraph, psb_fsg_core.h).  Shared by psb_search.cu and the emulation harness under tests/emul/.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "psb_fsg_core.h"

struct FsgFlat {
    std::vector<int32_t> buf;
    int P = 0, R = 0, L = 0, n_state = 0, n_null = 0, max_null = 0, CC = 0;
    size_t o_lp = 0, o_next = 0, o_sib = 0, o_ci = 0, o_leaf = 0, o_parent = 0, o_ctxt = 0, o_rlist = 0, o_rstate = 0,
           o_lto = 0, o_lall = 0, o_lnlp = 0, o_noff = 0, o_narc = 0;
    std::vector<int32_t> ssid, tmatid;
};

#define FSG_FAIL(...) do { char b_[256]; snprintf(b_, sizeof b_, __VA_ARGS__); err = b_; return -1; } while (0)

// pnodes [P][16] = ssid, tmatid, next, sibling, logs2prob, ci_ext, ppos, leaf, ctxt.bv[8];
// roots [n_state]; links [L][5] = from, to, wid, logs2prob, all-contexts; null arcs CSR by state.
static inline int
fsg_flatten(int P, const int32_t *pn, int n_state, const int32_t *roots, int L, const int32_t *links,
            const int32_t *nulloff, const int32_t *nullarc, int n_ci, FsgFlat &o, std::string &err)
{
    if (P <= 0 || n_state <= 0 || L < 0 || !pn || !roots || !nulloff) FSG_FAIL("fsg: empty graph");
    if (L > 0 && !links) FSG_FAIL("fsg: links missing");
    if (n_ci <= 0 || n_ci > 256) FSG_FAIL("fsg: %d CI phones (context bit vectors hold 256)", n_ci);
    o.P = P; o.L = L; o.n_state = n_state;
    std::vector<int32_t> parent((size_t)P, -1), rlist, rstate;
    std::vector<char> is_root((size_t)P, 0);
    for (int i = 0; i < P; ++i) {
        const int32_t *r = pn + (size_t)i * 16;
        if (r[3] < -1 || r[3] >= P) FSG_FAIL("fsg: pnode %d: sibling %d out of range", i, r[3]);
        if (r[5] < 0 || r[5] >= n_ci) FSG_FAIL("fsg: pnode %d: ci_ext %d out of range", i, r[5]);
        if (r[7]) { if (r[2] < 0 || r[2] >= L) FSG_FAIL("fsg: leaf %d: link %d out of range", i, r[2]); }
        else if (r[2] < -1 || r[2] >= P) FSG_FAIL("fsg: pnode %d: successor %d out of range", i, r[2]);
    }
    for (int l = 0; l < L; ++l) {
        if (links[l * 5 + 1] < 0 || links[l * 5 + 1] >= n_state) FSG_FAIL("fsg: link %d: destination state out of range", l);
    }
    long visited = 0;
    for (int s = 0; s < n_state; ++s) {
        if (roots[s] < -1 || roots[s] >= P) FSG_FAIL("fsg: root of state %d out of range", s);
        for (int p = roots[s]; p >= 0; p = pn[(size_t)p * 16 + 3]) {
            if (is_root[p] || ++visited > P) FSG_FAIL("fsg: root lists are not disjoint chains");
            is_root[p] = 1;
            rlist.push_back(p); rstate.push_back(s);
        }
    }
    visited = 0;
    for (int i = 0; i < P; ++i) {
        if (pn[(size_t)i * 16 + 7]) continue;
        for (int c = pn[(size_t)i * 16 + 2]; c >= 0; c = pn[(size_t)c * 16 + 3]) {
            if (++visited > P) FSG_FAIL("fsg: successor lists do not form a tree");
            if (is_root[c] || parent[c] >= 0) FSG_FAIL("fsg: pnode %d has more than one parent (the lextree must be a tree)", c);
            parent[c] = i;
        }
    }
    o.R = (int)rlist.size();
    o.n_null = nulloff[n_state];
    if (nulloff[0] != 0 || o.n_null < 0) FSG_FAIL("fsg: null-arc offsets must start at 0");
    o.max_null = 0;
    for (int s = 0; s < n_state; ++s) {
        const int n = nulloff[s + 1] - nulloff[s];
        if (n < 0) FSG_FAIL("fsg: null-arc offsets not monotone at state %d", s);
        if (n > o.max_null) o.max_null = n;
        for (int k = nulloff[s]; k < nulloff[s + 1]; ++k) {
            if (!nullarc || nullarc[k] < 0 || nullarc[k] >= L) FSG_FAIL("fsg: null arc %d: link out of range", k);
            if (links[nullarc[k] * 5 + 2] != -1) FSG_FAIL("fsg: null arc %d carries a word", k);
        }
    }
    o.CC = P * (1 + o.max_null) + 1;
    std::vector<int32_t> &b = o.buf;
    b.clear();
    auto col = [&](int c) { size_t at = b.size(); for (int i = 0; i < P; ++i) b.push_back(pn[(size_t)i * 16 + c]); return at; };
    o.o_lp = col(4); o.o_next = col(2); o.o_sib = col(3); o.o_ci = col(5); o.o_leaf = col(7);
    o.o_parent = b.size(); b.insert(b.end(), parent.begin(), parent.end());
    o.o_ctxt = b.size();
    for (int i = 0; i < P; ++i) for (int q = 0; q < 8; ++q) b.push_back(pn[(size_t)i * 16 + 8 + q]);
    o.o_rlist = b.size(); b.insert(b.end(), rlist.begin(), rlist.end());
    o.o_rstate = b.size(); b.insert(b.end(), rstate.begin(), rstate.end());
    o.o_lto = b.size(); for (int l = 0; l < L; ++l) b.push_back(links[l * 5 + 1]);
    o.o_lall = b.size(); for (int l = 0; l < L; ++l) b.push_back(links[l * 5 + 4]);
    o.o_lnlp = b.size(); for (int l = 0; l < L; ++l) b.push_back(links[l * 5 + 3] >> 10);      /* SENSCR_SHIFT */
    o.o_noff = b.size(); b.insert(b.end(), nulloff, nulloff + n_state + 1);
    o.o_narc = b.size(); if (o.n_null) b.insert(b.end(), nullarc, nullarc + o.n_null);
    b.push_back(0);
    o.ssid.resize(P); o.tmatid.resize(P);
    for (int i = 0; i < P; ++i) { o.ssid[i] = pn[(size_t)i * 16]; o.tmatid[i] = pn[(size_t)i * 16 + 1]; }
    return 0;
}

static inline void
fsg_graph_bind(const FsgFlat &o, const int32_t *base, FsgGraph &G)
{
    G.P = o.P; G.R = o.R; G.n_state = o.n_state; G.CC = o.CC;
    G.lp = base + o.o_lp; G.next = base + o.o_next; G.sib = base + o.o_sib; G.ci_ext = base + o.o_ci;
    G.leaf = base + o.o_leaf; G.parent = base + o.o_parent; G.ctxt = (const uint32_t *)(base + o.o_ctxt);
    G.root_list = base + o.o_rlist; G.root_state = base + o.o_rstate;
    G.link_to = base + o.o_lto; G.link_all = base + o.o_lall; G.link_nlp = base + o.o_lnlp;
    G.nulloff = base + o.o_noff; G.nullarc = base + o.o_narc;
}

// Per-utterance scratch, in int32 words, and its carving (the same on the host harness and the device).
static inline size_t
fsg_work_words(const FsgFlat &o, int n_emit)
{
    const size_t P = o.P, CC = o.CC, R = o.R > 0 ? o.R : 1;
    return (2 * (size_t)n_emit + 4) * P + 2 * P + 2 * P + 3 * (CC + 1) + 6 * CC + 16 * CC + 3 * CC + 8 * CC + 3 * (R + 1);
}
