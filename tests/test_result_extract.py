"""CPU: reading the hypothesis out of the search tables (psb_fsg_find_exit / psb_fsg_backtrace /
psb_ngram_find_exit / psb_ngram_backtrace: host code in the product library, no device work) against
the oracle's restatement on the reference's golden tables, and -- where oracle/_ref/libpsref.so is built
-- against the reference's own ps_get_hyp / ps_seg_iter on live decodes."""
import os

import numpy as np
import pytest

from conftest import golden
from oracle import oracle, refdrv
from pocketsphinx_b200 import api
from pocketsphinx_b200._lib import PsbError

FSG_TAGS = ("go", "go_hmmpf", "cmd", "cmd_wide", "cmd_hmmpf")
NG_TAGS = ("default", "wide", "narrow", "maxwpf", "abs", "pen", "lookahead", "flat_default", "flat_wide", "flat_narrow")


def _case(g, tag):
    return {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + ".")}


@pytest.mark.parametrize("tag", FSG_TAGS)
def test_fsg_exit_and_segments_on_the_golden_tables(tag):
    c = _case(golden("en_us_fsg.npz"), tag)
    hist, links, T, fin = c["hist"], c["links"], int(c["n_frames"]), int(c["final_state"])
    entry, score, seg = api.fsg_hyp(hist, links, T, fin)
    want = oracle.fsg_find_exit(hist, links, T, fin)
    assert (entry, score) == want and score == int(c["score"])
    vocab = str(c["vocab"]).split("\n")
    words = [vocab[w] for w in seg[:, 2] if w >= 0]
    assert [int(w) for w in seg[:, 2] if w >= 0] == oracle.fsg_hyp_wids(hist, links, entry)
    assert " ".join(w for w in words if not w.startswith("<") and not w.startswith("+")) == str(c["hyp"])
    # fsg_seg_bp2itor: segments tile the utterance, their scores add up to the exit's path score
    assert seg[0, 3] == 0 and seg[-1, 0] == entry and seg[-1, 4] == hist[entry, 1]
    for a, b in zip(seg[:-1], seg[1:]):
        assert b[3] == min(a[4] + 1, b[4])
    assert int(seg[:, 5].sum() + seg[:, 6].sum()) == score
    # every frame, partial results (final = 0), as ps_get_hyp serves them mid-utterance
    for f in range(0, T, 7):
        e, s, sg = api.fsg_hyp(hist, links, f, fin, final=False)
        we, ws = oracle.fsg_find_exit(hist, links, f, fin, final=False)
        assert e == we and (e <= 0 or s == ws)
        assert len(sg) == (0 if e <= 0 else len(oracle.fsg_hyp_wids(hist, links, e)) + int((links[hist[sg[:, 0], 0], 2] < 0).sum()))


@pytest.mark.parametrize("tag", NG_TAGS)
def test_ngram_exit_and_chain_on_the_golden_tables(tag):
    g = golden("en_us_fwdtree.npz")
    c = _case(g, tag)
    d = _case(g, "default")
    bp, bp_idx = c["bp"], c["bp_idx"]
    info = c["info"] if "info" in c else d["info"]
    words = c["words"] if "words" in c else d["words"]
    vocab = str(c["vocab"] if "vocab" in c else d["vocab"]).split("\n")
    T = len(bp_idx) - 1
    entry, score, seg = api.ngram_hyp(bp, bp_idx, T, int(info[20]))
    assert (entry, score) == oracle.fwdtree_find_exit(bp, bp_idx, T, int(info[20]))
    hyp = " ".join(vocab[int(words[w][5])] for w in seg[:, 1] if not words[w][4] and int(words[w][5]) not in (int(info[19]), int(info[20])))
    assert hyp == oracle.fwdtree_hyp(bp, entry, words, vocab, int(info[19]), int(info[20]))
    if "hyp" in c and "score" in c:
        assert hyp == str(c["hyp"]) and score == int(c["score"])
    assert seg[0, 2] == 0 and seg[0, 1] == int(info[19]) and seg[-1, 0] == entry
    assert (seg[1:, 2] == seg[:-1, 3] + 1).all() and (seg[:, 4] == bp[seg[:, 0], 4]).all()


def test_tables_without_a_hypothesis():
    c = _case(golden("en_us_fsg.npz"), "go")
    hist, links, fin = c["hist"], c["links"], int(c["final_state"])
    assert api.fsg_hyp(hist[:1], links, 10, fin)[0] == 0                       # only the start entry
    assert api.fsg_hyp(hist, links, -1, fin)[0] == 0                           # before the first frame
    assert api.fsg_hyp(hist, links, int(c["n_frames"]), 10 ** 6)[0] == -1      # the final state was not reached
    d = _case(golden("en_us_fwdtree.npz"), "default")
    assert api.ngram_hyp(d["bp"], d["bp_idx"], 0, 1)[0] == -1
    idx = np.zeros(11, np.int32)
    assert api.ngram_hyp(d["bp"], idx, 10, 1)[0] == -1                         # no frame has exits


def test_malformed_tables_are_errors_not_walks():
    c = _case(golden("en_us_fsg.npz"), "go")
    hist, links, fin = c["hist"].copy(), c["links"], int(c["final_state"])
    entry = api.fsg_hyp(hist, links, int(c["n_frames"]), fin)[0]
    hist[entry, 3] = entry                                                     # a chain that does not go backwards
    with pytest.raises(PsbError):
        api.fsg_hyp(hist, links, int(c["n_frames"]), fin)
    hist = c["hist"].copy()
    hist[entry, 0] = len(links)
    with pytest.raises(PsbError):
        api.fsg_hyp(hist, links, int(c["n_frames"]), fin)
    d = _case(golden("en_us_fwdtree.npz"), "default")
    bp = d["bp"].copy()
    e = api.ngram_hyp(bp, d["bp_idx"], len(d["bp_idx"]) - 1, int(d["info"][20]))[0]
    bp[e, 3] = e + 1
    with pytest.raises(PsbError):
        api.ngram_hyp(bp, d["bp_idx"], len(d["bp_idx"]) - 1, int(d["info"][20]))
    idx = d["bp_idx"].copy()
    idx[-2] = len(bp) + 5
    idx[-1] = len(bp) + 9
    with pytest.raises(PsbError):
        api.ngram_hyp(d["bp"], idx, len(idx) - 1, int(d["info"][20]))


live = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")


@live
@pytest.mark.parametrize("which", ["commands", "tidigits"])
def test_fsg_segments_match_the_reference_iterator(which):
    rd = os.path.dirname(refdrv.LIB_PATH)
    pcm = np.fromfile(os.path.join(rd, "data", "goforward.raw"), np.int16)
    if which == "commands":
        r = refdrv.fsg(os.path.join(rd, "model", "en-us"), os.path.join(rd, "model", "cmudict-en-us.dict"),
                       os.path.join(os.path.dirname(__file__), "golden", "commands.fsg"), pcm)
    else:
        r = refdrv.fsg(os.path.join(rd, "model", "tidigits_hmm"), os.path.join(rd, "model", "tidigits_lm", "tidigits.dic"),
                       os.path.join(rd, "model", "tidigits_lm", "tidigits.fsg"), pcm)
    entry, score, seg = api.fsg_hyp(r["hist"], r["links"], r["n_frames"], r["final_state"])
    assert score == r["score"] and len(seg) == len(r["seg"]) > 0
    for s, (word, sf, ef, ascr, lscr) in zip(seg, r["seg"]):
        assert (r["vocab"][s[2]] if s[2] >= 0 else "(NULL)") == word
        assert (int(s[3]), int(s[4]), int(s[5]), int(s[6])) == (int(sf), int(ef), int(ascr), int(lscr))


@live
@pytest.mark.parametrize("kv", [dict(), dict(fwdflat="yes"), dict(beam="1e-60", wbeam="1e-40", maxwpf="5")])
def test_ngram_segments_match_the_reference_iterator(kv):
    rd = os.path.dirname(refdrv.LIB_PATH)
    hd, lm, dic = os.path.join(rd, "model", "en-us"), os.path.join(rd, "data", "turtle.lm.bin"), os.path.join(rd, "data", "turtle.dic")
    pcm = np.fromfile(os.path.join(rd, "data", "goforward.raw"), np.int16)
    r = refdrv.fwdtree(hd, lm, dic, pcm, **kv)
    full = refdrv.decode(hd, lm, dic, pcm, bestpath="no", compallsen="yes", pl_window="0", **dict(dict(fwdflat="no"), **kv))
    entry, score, seg = api.ngram_hyp(r["bp"], r["bp_idx"], r["n_frame"], r["finish_wid"])
    lines = [l.split() for l in full["seg"].split("\n") if l]
    assert score == full["score"] and len(lines) == len(seg) > 0
    for s, (word, sf, ef, _, _) in zip(seg, lines):
        assert (int(s[2]), int(s[3])) == (int(sf), int(ef))
        assert r["vocab"][int(r["words"][s[1]][5])] == word.split("(")[0]


@live
@pytest.mark.parametrize("kv", [dict(), dict(fwdflat="yes"), dict(fwdflat="yes", fwdflatlw="7.5", lw="5"),
                                dict(beam="1e-60", wbeam="1e-40", maxwpf="5"), dict(silprob="0.02", fillprob="1e-5")])
@pytest.mark.parametrize("arrays", [False, True])
def test_ngram_segment_scores_match_the_reference_iterator(kv, arrays):
    """ascr / lscr of every segment (ngram_search_bp2itor): right-context exit scores from the score stack,
    LM scores from the dense table or the LM arrays, scaled by the float32 fwdflatlw / lw after a second pass."""
    rd = os.path.dirname(refdrv.LIB_PATH)
    hd, lm, dic = os.path.join(rd, "model", "en-us"), os.path.join(rd, "data", "turtle.lm.bin"), os.path.join(rd, "data", "turtle.dic")
    pcm = np.fromfile(os.path.join(rd, "data", "goforward.raw"), np.int16)
    r = refdrv.fwdtree(hd, lm, dic, pcm, dense_lm=not arrays, **kv)
    lma = refdrv.lm_arrays(hd, lm, dic, **kv)[0] if arrays else None
    full = refdrv.decode(hd, lm, dic, pcm, bestpath="no", compallsen="yes", pl_window="0", **dict(dict(fwdflat="no"), **kv))
    entry, score, chain = api.ngram_hyp(r["bp"], r["bp_idx"], r["n_frame"], r["finish_wid"])
    seg = api.ngram_segments(r["info"], r["model"], r["bp"], r["bss"], entry, lm_arrays=lma, second_pass=kv.get("fwdflat") == "yes")
    assert np.array_equal(seg[:, :5], chain)
    lines = [l.split() for l in full["seg"].split("\n") if l]
    assert len(lines) == len(seg) > 3
    for s, (word, sf, ef, ascr, lscr) in zip(seg, lines):
        assert (int(s[2]), int(s[3]), int(s[5]), int(s[6])) == (int(sf), int(ef), int(ascr), int(lscr)), word
    assert any(int(l[4]) != 0 for l in lines)


def test_ngram_segments_reject_tables_that_leave_the_model():
    g = golden("en_us_fwdtree.npz")
    d = _case(g, "default")
    e = api.ngram_hyp(d["bp"], d["bp_idx"], len(d["bp_idx"]) - 1, int(d["info"][20]))[0]
    seg = api.ngram_segments(d["info"], d["model"], d["bp"], d["bss"], e)
    assert len(seg) > 3 and int(seg[:, 5].sum() + seg[:, 6].sum()) == int(d["bp"][e, 4]) + sum(
        int(d["bp"][p, 4]) - _start(d, p, int(d["bp"][b, 2])) for b, p in zip(seg[1:, 0], seg[:-1, 0]))
    with pytest.raises(PsbError):
        api.ngram_segments(d["info"], d["model"][:1000], d["bp"], d["bss"], e)
    with pytest.raises(PsbError):
        api.ngram_segments(d["info"], d["model"], d["bp"], d["bss"][:4], e)
    bp = d["bp"].copy()
    bp[seg[1, 0], 6] = 10 ** 6
    with pytest.raises(PsbError):
        api.ngram_segments(d["info"], d["model"], bp, d["bss"], e)
    bp = d["bp"].copy()
    bp[seg[1, 0], 8] = 10 ** 6
    with pytest.raises(PsbError):
        api.ngram_segments(d["info"], d["model"], bp, d["bss"], e)


def _start(d, p, wid):
    """ngram_search_exit_score of entry p into word wid, from the oracle's view of the tables."""
    info, model, bp, bss = d["info"], d["model"], d["bp"], d["bss"]
    nc = int(info[6])
    o_words = int(info[2]) * 5 + int(info[3]) * 6
    words = model[o_words:o_words + int(info[1]) * 8].reshape(-1, 8)
    o_cimap = o_words + int(info[1]) * 8 + int(info[4]) * 5 + nc * nc + nc ** 3
    cimap = model[o_cimap:o_cimap + nc ** 3].reshape(nc, nc, nc)
    if bp[p, 9] == -1:
        return int(bp[p, 4])
    return int(bss[bp[p, 5] + cimap[bp[p, 8], bp[p, 9], words[wid, 0]]])
