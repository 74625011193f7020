"""Cross-word triphone tables (dict2pid.c) from a model definition and a dictionary, without the reference:
the right-context fan-out of every (last phone, second-last phone) pair, compressed to its distinct senone
sequences (rssid: n_ssid, ssid[], cimap[]), and the left-context senone sequences of every (first phone,
second phone) pair (ldiph_lc) -- the `rs_n | rs_ssid | rs_cimap | ldiph_lc` sections of the flattened n-gram
search (include/psb200.h, psb_ngram_desc_t.model).  tests/test_dict2pid.py compares them with the sections
the maintainer-side binding exports from the reference's own dict2pid_t."""
import numpy as np

BAD_SSID = 0xFFFF
WPOS_INTERNAL, WPOS_BEGIN, WPOS_END, WPOS_SINGLE = 0, 1, 2, 3


class TriphoneIndex:
    """bin_mdef_phone_id / bin_mdef_phone_id_nearest (bin_mdef.c:744-857) on the arrays s3io.read_mdef returns."""

    def __init__(self, md):
        self.n_ci = md["n_ciphone"]
        self.sil = md["sil"]
        self.filler = [bool(x) for x in md["phone_filler"]]
        self.ssid = md["phone_ssid"]
        t = md["cd_tree"]
        self.tree = None if t is None else (t["ctx"].tolist(), t["n_down"].tolist(), t["down"].tolist())
        self._kids = {}

    def _find(self, start, count, ctx):
        key = (start, count)
        kids = self._kids.get(key)
        if kids is None:
            kids = {}
            for i in range(start, start + count):
                kids.setdefault(self.tree[0][i], i)              # the first match wins, as in the linear scan
            self._kids[key] = kids
        return kids.get(ctx, -1)

    def phone_id(self, ci, lc, rc, wpos):
        if lc < 0 or rc < 0:
            return ci
        if self.tree is None:
            return -1
        ctx = (wpos, ci, self.sil if (self.sil >= 0 and self.filler[lc]) else lc,
               self.sil if (self.sil >= 0 and self.filler[rc]) else rc)
        start, count = 0, 4
        for level in range(4):
            i = self._find(start, count, ctx[level])
            if i < 0:
                return -1
            if self.tree[1][i] == 0:
                return self.tree[2][i]
            start, count = self.tree[2][i], self.tree[1][i]
        return -1

    def nearest(self, b, l, r, pos):
        if l < 0 or r < 0:
            return b
        order = [pos] + [p for p in range(4) if p != pos]
        for p in order:
            pid = self.phone_id(b, l, r, p)
            if pid >= 0:
                return pid
        if self.sil >= 0:
            nl = self.sil if (self.filler[l] or pos in (WPOS_BEGIN, WPOS_SINGLE)) else l
            nr = self.sil if (self.filler[r] or pos in (WPOS_END, WPOS_SINGLE)) else r
            if nl != l or nr != r:
                for p in order:
                    pid = self.phone_id(b, nl, nr, p)
                    if pid >= 0:
                        return pid
        return b


def _compress(row, n):
    """compress_table (dict2pid.c:52-80): the distinct entries in order of first appearance, and for every
    right context the index of its entry."""
    com, cimap = [BAD_SSID] * n, [-1] * n
    for r in range(n):
        t = 0
        found = False
        while t < r and com[t] != BAD_SSID:
            if row[r] == com[t]:
                found = True
                break
            t += 1
        if not found:
            com[t] = row[r]
        cimap[r] = t
    k = 0
    while k < n and com[k] != BAD_SSID:
        k += 1
    return com, cimap, k


def build(md, prons):
    """dict2pid_build (dict2pid.c:391-499) over the dictionary's pronunciations in word-id order (lists of
    CI phone ids).  Returns rs_n [n][n], rs_ssid [n][n][n] (-1 past n_ssid), rs_cimap [n][n][n] (-1 where a
    pair never ends a word), ldiph_lc [n][n][n] (0xffff where a pair never starts one)."""
    n = md["n_ciphone"]
    tri = TriphoneIndex(md)
    ssid = md["phone_ssid"]
    ldiph = np.full((n, n, n), BAD_SSID, np.int32)
    rdiph = np.full((n, n, n), BAD_SSID, np.int32)
    lrdiph = np.full((n, n, n), BAD_SSID, np.int32)
    seen_l, seen_r, single = set(), set(), set()
    for p in prons:
        if len(p) >= 2:
            b, r = p[0], p[1]
            if (b, r) not in seen_l:
                seen_l.add((b, r))
                for l in range(n):
                    ldiph[b, r, l] = ssid[tri.nearest(b, l, r, WPOS_BEGIN)]
            l, b = p[-2], p[-1]
            if (b, l) not in seen_r:
                seen_r.add((b, l))
                for r in range(n):
                    rdiph[b, l, r] = ssid[tri.nearest(b, l, r, WPOS_END)]
        elif len(p) == 1 and p[0] not in single:
            b = p[0]
            single.add(b)
            for l in range(n):                                  # populate_lrdiph (:270-298): also the silence-context
                for r in range(n):                              # rows of the other two tables, unconditionally
                    s = ssid[tri.nearest(b, l, r, WPOS_SINGLE)]
                    lrdiph[b, l, r] = s
                    if r == tri.sil:
                        ldiph[b, r, l] = s
                    if l == tri.sil:
                        rdiph[b, l, r] = s
    rs_n = np.zeros((n, n), np.int32)
    rs_ssid = np.full((n, n, n), -1, np.int32)
    rs_cimap = np.full((n, n, n), -1, np.int32)
    for b in range(n):
        for l in range(n):
            com, cimap, k = _compress(rdiph[b, l].tolist(), n)
            if com[0] != BAD_SSID:
                rs_n[b, l] = k
                rs_ssid[b, l, :k] = com[:k]
                rs_cimap[b, l] = cimap
    return dict(rs_n=rs_n, rs_ssid=rs_ssid, rs_cimap=rs_cimap, ldiph_lc=ldiph, lrdiph_rc=lrdiph)


def alignment_phones(md, prons, tabs, wids):
    """ps_alignment_populate (ps_alignment.c:131-218): the phone chain of a word sequence with its cross-word
    contexts -- silence to the left of the first word and to the right of the last, the neighbours' last / first
    phones in between -- as (ssid, tmatid, cipid) per phone: what HmmContext.align takes.  prons: the dictionary's
    pronunciations (CI phone ids by word id), tabs: build(md, prons), wids: the transcript's dictionary ids."""
    tri = TriphoneIndex(md)
    ssid_of, tmat_of, sil = md["phone_ssid"], md["phone_tmat"], md["sil"]
    ssid, cipid = [], []
    lc = sil
    for i, w in enumerate(wids):
        p = prons[w]
        rc = prons[wids[i + 1]][0] if i + 1 < len(wids) else sil
        if len(p) == 1:
            ssid.append(int(tabs["lrdiph_rc"][p[0], lc, rc]))
        else:
            ssid.append(int(tabs["ldiph_lc"][p[0], p[1], lc]))
        cipid.append(p[0])
        for j in range(1, len(p) - 1):
            ssid.append(int(ssid_of[tri.nearest(p[j], p[j - 1], p[j + 1], WPOS_INTERNAL)]))
            cipid.append(p[j])
        if len(p) > 1:
            ssid.append(int(tabs["rs_ssid"][p[-1], p[-2], tabs["rs_cimap"][p[-1], p[-2], rc]]))
            cipid.append(p[-1])
        lc = p[-1]
    if BAD_SSID in ssid or -1 in ssid:
        raise ValueError("a word of the transcript has no senone sequence in this context")
    cipid = np.array(cipid, np.int32)
    return np.array(ssid, np.int32), tmat_of[cipid].astype(np.int32), cipid


def keyphrase_phones(md, prons, tabs, wids):
    """The HMM chain kws_search_reinit builds for one keyphrase (kws_search.c:552-585): every word with silence as
    its outer context (first phone: ldiph_lc with a silence left context -- also for single-phone words, whose
    right context is silence too; last phone: the right-context entry for silence; word-internal triphones in
    between) -> (ssid, tmatid) per phone, what HmmContext.kws takes per keyphrase."""
    tri = TriphoneIndex(md)
    ssid_of, tmat_of, sil = md["phone_ssid"], md["phone_tmat"], md["sil"]
    ssid, cipid = [], []
    for w in wids:
        p = prons[w]
        for j, ci in enumerate(p):
            if j == 0:
                ssid.append(int(tabs["ldiph_lc"][ci, p[1] if len(p) > 1 else sil, sil]))
            elif j == len(p) - 1:
                ssid.append(int(tabs["rs_ssid"][ci, p[j - 1], tabs["rs_cimap"][ci, p[j - 1], sil]]))
            else:
                ssid.append(int(ssid_of[tri.nearest(ci, p[j - 1], p[j + 1], WPOS_INTERNAL)]))
            cipid.append(ci)
    cipid = np.array(cipid, np.int32)
    return np.array(ssid, np.int32), tmat_of[cipid].astype(np.int32)
