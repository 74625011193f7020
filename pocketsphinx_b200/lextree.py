"""The flattened n-gram search (psb_ngram_desc_t: info[40] + model sections, include/psb200.h) built from the
files alone -- model definition, dictionary, LM word list and the search settings -- instead of being exported
from a reference decoder (integration/ps_search_cuda.c:cuda_ngram_export).  Restates, with their orderings,
init_search_tree / create_search_channels (ngram_search_fwdtree.c:66-330: root channels per first diphone in
order of first use, interior channels found by senone sequence along the `alt` chain and appended at its
end, words hanging off their penultimate channel in homophone chains, single-phone words outside the tree)
and ngram_search_calc_beams (ngram_search.c:92-132).  tests/test_lextree.py compares info and every model
section with the binding's export on the reference's own models."""
import math

import numpy as np

from . import dict2pid as d2p

DEFAULTS = dict(beam="1e-48", wbeam="7e-29", pbeam="1e-48", lpbeam="1e-40", lponlybeam="7e-29", fwdflatbeam="1e-64",
                fwdflatwbeam="7e-29", maxwpf="-1", maxhmmpf="30000", wip="0.65", nwpen="1.0", pip="1.0", silprob="0.005",
                fillprob="1e-8", lw="6.5", fwdflatlw="8.5", fwdflatefwid="4", fwdflatsfwin="25", fwdflat="yes", logbase="1.0001")


def _logs(p, logbase):
    """logmath_log(lmath, p) >> SENSCR_SHIFT (util/logmath.c, shift 0)."""
    v = -(1 << 31) >> 2 if p <= 0 else int(math.log(p) * (1.0 / math.log(logbase)))
    return v >> 10


def build_ngram_search(md, words, prons, base, filler_start, widmap, unknown_wid=-1, **config):
    """md: s3io.read_mdef; words / prons: strings and CI phone ids of every dictionary word in id order, base /
    filler_start: lmio.read_dict; widmap: the LM word id of every dictionary word (lmio.lm_arrays' widmap) and the LM's
    <UNK> id (or -1).  Returns (info int32[40], model int32[...]) without the dense trigram table: scores come
    from the LM arrays."""
    cfg = dict(DEFAULTS)
    cfg.update({k: str(v) for k, v in config.items()})
    lb = float(cfg["logbase"])
    n_words, n_ci, sil = len(prons), md["n_ciphone"], md["sil"]
    ssid_of, tmat_of = md["phone_ssid"], md["phone_tmat"]
    tri = d2p.TriphoneIndex(md)
    start_wid, finish_wid, silence_wid = words.index("<s>"), words.index("</s>"), words.index("<sil>")   # dict.c:383-385
    filler_end = n_words - 1

    def is_filler(w):                                         # dict_filler_word (dict.c:411-422): <s> and </s> are not
        b = int(base[w])
        return b not in (start_wid, finish_wid) and filler_start <= b <= filler_end
    known = [int(widmap[base[w]]) != unknown_wid for w in range(n_words)]          # ngram_model_set_known_wid
    single = [len(p) == 1 for p in prons]

    def internal(w, pos):                                     # dict2pid_internal (dict2pid.c:370-388)
        p = prons[w]
        return int(ssid_of[tri.nearest(p[pos], p[pos - 1], p[pos + 1], d2p.WPOS_INTERNAL)])
    # create_search_channels
    homophone = [-1] * n_words
    tail = {}
    roots, root_of = [], {}                                   # [ciphone, ci2phone, penult_wid, first child, tmatid]
    nodes = []                                                # [ssid, tmatid, ciphone, penult_wid, next, alt]
    kids = {}                                                 # parent key -> {ssid: node}, and the chain's last node
    single_lm = []

    def hang(get, put, w):
        j = get()
        if j < 0:
            put(w)
        else:
            t = tail.get(j, j)
            while homophone[t] >= 0:
                t = homophone[t]
            homophone[t] = w
            tail[j] = w

    def child(parent_key, set_first, ssid, ci):
        k = kids.get(parent_key)
        if k is None:
            k = kids[parent_key] = [{}, -1]
        n = k[0].get(ssid)
        if n is None:
            n = len(nodes)
            nodes.append([ssid, int(tmat_of[ci]), ci, -1, -1, -1])
            if k[1] < 0:
                set_first(n)
            else:
                nodes[k[1]][5] = n                             # appended at the end of the alt chain
            k[0][ssid] = n
            k[1] = n
        return n
    for w in range(n_words):
        if not known[w]:
            continue
        if single[w]:
            single_lm.append(w)
            continue
        p = prons[w]
        key = (p[0], p[1])
        r = root_of.get(key)
        if r is None:
            r = root_of[key] = len(roots)
            roots.append([p[0], p[1], -1, -1, int(tmat_of[p[0]])])
        if len(p) == 2:
            hang(lambda: roots[r][2], lambda v: roots[r].__setitem__(2, v), w)
        else:
            n = child(("r", r), lambda v: roots[r].__setitem__(3, v), internal(w, 1), p[1])
            for pos in range(2, len(p) - 1):
                prev = n
                n = child(("n", prev), lambda v, prev=prev: nodes[prev].__setitem__(4, v), internal(w, pos), p[pos])
            hang(lambda: nodes[n][3], lambda v: nodes[n].__setitem__(3, v), w)
    n_1ph_lm = len(single_lm)
    single_wid = list(single_lm)
    for w in range(n_words):                                  # fillers that the LM does not know
        real = not is_filler(w) and int(base[w]) not in (start_wid, finish_wid)   # dict_real_word (dict.c:431-442)
        if single[w] and not real and not known[w]:
            single_wid.append(w)
    # the binding numbers interior channels depth-first: a channel, its subtree, then its alternatives
    order = []
    for r in roots:
        stack = [r[3]]
        while stack:
            h = stack.pop()
            if h < 0:
                continue
            order.append(h)
            stack.append(nodes[h][5])                          # alternatives after the subtree ...
            stack.append(nodes[h][4])                          # ... which comes first
    new_id = {h: i for i, h in enumerate(order)}
    new_id[-1] = -1
    out = []
    for r in roots:
        out += [r[0], r[1], r[2], new_id[r[3]], r[4]]
    for h in order:
        c = nodes[h]
        out += [c[0], c[1], c[2], c[3], new_id[c[4]], new_id[c[5]]]
    for w in range(n_words):
        p = prons[w]
        out += [p[0], p[-1], -1 if single[w] else p[-2], int(single[w]), int(is_filler(w)), int(base[w]),
                homophone[w], -1]
    out += single_wid
    for w in single_wid:
        ci = prons[w][0]
        out += [ci, sil, int(ssid_of[ci]), int(tmat_of[ci])]
    tabs = d2p.build(md, prons)
    model = [np.array(out, np.int32), tabs["rs_n"].ravel(), tabs["rs_ssid"].ravel(), tabs["rs_cimap"].ravel(), tabs["ldiph_lc"].ravel()]
    # second pass: LM membership, pronunciations and their word-internal senone sequences
    n_pron = sum(len(p) for p in prons)
    off = np.zeros(n_words + 1, np.int32)
    off[1:] = np.cumsum([len(p) for p in prons])
    flat = np.array([x for p in prons for x in p], np.int32)
    inner = np.full(n_pron, -1, np.int32)
    for w in range(n_words):
        for j in range(1, len(prons[w]) - 1):
            inner[off[w] + j] = internal(w, j)
    model += [np.array(known, np.int32), off, flat, inner]
    info = np.zeros(40, np.int32)
    info[1:8] = (n_words, len(roots), len(order), len(single_wid), n_1ph_lm, n_ci, sil)
    for i, k in ((8, "beam"), (9, "pbeam"), (10, "wbeam"), (11, "lpbeam"), (12, "lponlybeam")):
        info[i] = _logs(float(cfg[k]), lb)
    info[13], info[14] = int(cfg["maxhmmpf"]), int(cfg["maxwpf"])
    info[15], info[16] = _logs(float(cfg["nwpen"]), lb), _logs(float(cfg["pip"]), lb)
    info[17] = info[16] + _logs(float(cfg["silprob"]), lb)
    info[18] = info[16] + _logs(float(cfg["fillprob"]), lb)
    info[19], info[20], info[21] = start_wid, finish_wid, silence_wid
    info[22], info[23] = filler_start, filler_end
    info[28], info[29] = _logs(float(cfg["fwdflatbeam"]), lb), _logs(float(cfg["fwdflatwbeam"]), lb)
    if cfg["fwdflat"] in ("yes", "1", "true", "True"):
        info[30], info[31] = int(cfg["fwdflatefwid"]), int(cfg["fwdflatsfwin"])
    info[32] = np.array([float(cfg["fwdflatlw"]) / float(cfg["lw"])], np.float32).view(np.int32)[0]
    info[33] = n_pron
    return info, np.concatenate(model).astype(np.int32)


def ngram_search_from_files(hmm_dir, dict_file, lm_file, filler_dict=None, **config):
    """Everything `HmmContext.ngram_fwdtree / ngram_fwdflat / ngram_two_pass` take besides the senone scores,
    from an acoustic-model directory, a dictionary and a binary trie LM (the reference's -hmm / -dict / -lm):
    dict(info, model, lm_arrays, ci_tmat, ci_ssid, words, base).  filler_dict defaults to the model's `noisedict`
    (-fdict); config: the reference's search settings by name (beam, wbeam, lw, wip, fwdflatlw, ...)."""
    import os

    from . import lmio, s3io
    md = s3io.read_mdef(os.path.join(hmm_dir, "mdef"))
    if filler_dict is None:
        nd = os.path.join(hmm_dir, "noisedict")
        filler_dict = nd if os.path.exists(nd) else None
    words, prons, base, filler_start = lmio.read_dict(dict_file, filler_dict, md["ciname"])
    ci = {n: i for i, n in enumerate(md["ciname"])}
    lm = lmio.read_lm_bin(lm_file)
    if "</s>" not in lm["words"]:
        raise ValueError("%s: the language model does not contain </s>" % lm_file)        # ngram_search.c:196-202
    arr = lmio.lm_arrays(lm, words, lw=float(config.get("lw", DEFAULTS["lw"])), wip=float(config.get("wip", DEFAULTS["wip"])),
                         logbase=float(config.get("logbase", DEFAULTS["logbase"])))
    unk = lm["words"].index("<UNK>") if "<UNK>" in lm["words"] else -1
    info, model = build_ngram_search(md, words, [[ci[x] for x in p] for p in prons], base, filler_start,
                                     arr[10:10 + len(words)], unk, **config)
    n_ci = md["n_ciphone"]
    return dict(info=info, model=model, lm_arrays=arr, ci_tmat=md["phone_tmat"][:n_ci].astype(np.int32),
                ci_ssid=md["phone_ssid"][:n_ci].astype(np.int32), words=words, base=base)
