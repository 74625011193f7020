"""CPU: the C restatement of the first decoding pass -- ngram_search_fwdtree.c plus the
backpointer-table half of ngram_search.c (oracle/ps_oracle.c: pso_fwdtree_run) -- against what the
reference's own search produced on goforward.raw with the turtle LM (tests/golden/en_us_fwdtree.npz):
every bp_table entry (frame, valid, wid, bp, score, s_idx, real_wid, prev_real_wid, last phones), the
whole right-context score stack, bp_table_idx, and the exit / hypothesis ngram_search_find_exit and
ngram_search_bp_hyp derive from them, on the reference's senone scores."""
import numpy as np
import pytest

from conftest import golden

TAGS = ("default", "wide", "narrow", "maxwpf", "abs", "pen", "lookahead")


def _case(g, tag):
    return {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + ".")}


@pytest.mark.parametrize("tag", TAGS)
def test_fwdtree_oracle_matches_reference_golden(tag):
    from oracle import oracle
    m = golden("en_us_ptm_model.npz")
    scr = golden("en_us_goforward.npz")["senscr"]
    c = _case(golden("en_us_fwdtree.npz"), tag)
    n_ci = int(c["info"][6])
    gf = golden("en_us_goforward.npz")
    la = dict(pl_pen=gf["pl_pen"], pl_window=int(gf["pl_params"][4])) if tag == "lookahead" else {}
    bp, bss, bp_idx = oracle.fwdtree_run(m["tp"], m["sseq"], m["phone_tmat"][:n_ci], c["info"], c["model"], scr, **la)
    assert bp.shape == c["bp"].shape and np.array_equal(bp, c["bp"])
    assert np.array_equal(bss, c["bss"]) and np.array_equal(bp_idx, c["bp_idx"])
    b, score = oracle.fwdtree_find_exit(bp, bp_idx, len(scr), int(c["info"][20]))
    assert b >= 0 and score == int(c["score"])
    vocab = str(c["vocab"]).split("\n")
    assert oracle.fwdtree_hyp(bp, b, c["words"], vocab, int(c["info"][19]), int(c["info"][20])) == str(c["hyp"])


def test_fwdtree_golden_covers_the_interesting_paths():
    g = golden("en_us_fwdtree.npz")
    d, wide, mw, ab = (_case(g, t) for t in ("default", "wide", "maxwpf", "abs"))
    assert len(wide["bp"]) > 4 * len(d["bp"])
    assert (mw["bp"][:, 1] == 0).sum() > 100 and (d["bp"][:, 1] == 0).sum() == 0      # -maxwpf invalidated exits
    assert len(ab["bp"]) < len(d["bp"])                                               # histogram pruning narrowed the beam
    multi = d["bp"][:, 5] >= 0
    assert multi.sum() > 300 and (d["bss"] > -0x20000000).sum() > 10 * multi.sum()        # several right contexts per exit
    assert (d["bp"][:, 6] != d["bp"][:, 2]).any()                                     # fillers inherit the LM state


@pytest.mark.parametrize("tag", ("flat_default", "flat_wide", "flat_narrow"))
def test_fwdflat_oracle_matches_reference_golden(tag):
    """Both passes chained: pso_fwdtree_run's table feeds pso_fwdflat_run (ngram_search_fwdflat.c), whose
    table must equal the one the reference's second pass left behind."""
    from oracle import oracle
    m = golden("en_us_ptm_model.npz")
    gf = golden("en_us_goforward.npz")
    scr = gf["senscr"]
    c = _case(golden("en_us_fwdtree.npz"), tag)
    n_ci = int(c["info"][6])
    la = dict(pl_pen=gf["pl_pen"], pl_window=int(gf["pl_params"][4])) if tag == "flat_default" else {}
    bp1, _, _ = oracle.fwdtree_run(m["tp"], m["sseq"], m["phone_tmat"][:n_ci], c["info"], c["model"], scr, **la)
    bp, bss, bp_idx = oracle.fwdflat_run(m["tp"], m["sseq"], m["phone_tmat"][:n_ci], m["phone_ssid"][:n_ci], c["info"],
                                         c["model"], bp1, scr)
    assert bp.shape == c["bp"].shape and np.array_equal(bp, c["bp"])
    assert np.array_equal(bss, c["bss"]) and np.array_equal(bp_idx, c["bp_idx"])
    b, score = oracle.fwdtree_find_exit(bp, bp_idx, len(scr), int(c["info"][20]))
    assert b >= 0 and score == int(c["score"])
    vocab = str(c["vocab"]).split("\n")
    assert oracle.fwdtree_hyp(bp, b, c["words"], vocab, int(c["info"][19]), int(c["info"][20])) == str(c["hyp"])
    if tag == "flat_default":                                  # the first pass's table with look-ahead is the fixture's, too
        assert np.array_equal(bp1, _case(golden("en_us_fwdtree.npz"), "lookahead")["bp"])
