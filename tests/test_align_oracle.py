"""CPU: the C restatement of state_align_search.c (oracle/ps_oracle.c: pso_align_run) against the
alignments the reference's own search produced (tests/golden/en_us_align.npz), on the reference's
senone scores; and, in the build container, against the reference run live."""
import os

import numpy as np
import pytest

from conftest import golden


def _check(tag, g, m, scr):
    from oracle import oracle
    rc, st, du, sc = oracle.align_run(m["tp"], m["sseq"], g[tag + "_ssid"], g[tag + "_tmatid"], scr)
    assert rc == 0
    assert np.array_equal(st, g[tag + "_start"]) and np.array_equal(du, g[tag + "_dur"]), tag
    assert np.array_equal(sc, g[tag + "_score"]), tag
    assert du.sum() == len(scr) and st[0] == 0


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_align_oracle_matches_reference_golden(tag):
    _check(tag, golden("en_us_align.npz"), golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")["senscr"])


def test_align_oracle_failure_modes():
    from oracle import oracle
    g, m = golden("en_us_align.npz"), golden("en_us_ptm_model.npz")
    scr = golden("en_us_goforward.npz")["senscr"]
    # fewer frames than emitting states: the final state is never reached
    rc, st, du, sc = oracle.align_run(m["tp"], m["sseq"], g["a_ssid"], g["a_tmatid"], scr[:20])
    assert rc == -1 and (st == -1).all()
    # the first phone must end before its successor may start: nothing survives
    n = len(g["a_ssid"])
    sf, ef = np.zeros(n, np.int32), np.full(n, 2**31 - 1, np.int32)
    ef[0], sf[1] = 5, 100
    rc, st, du, sc = oracle.align_run(m["tp"], m["sseq"], g["a_ssid"], g["a_tmatid"], scr, sf=sf, ef=ef)
    assert rc == -1 and (st == -1).all()
    # constraints that only delay a phone change the segmentation but still cover every frame
    sf, ef = np.zeros(n, np.int32), np.full(n, 2**31 - 1, np.int32)
    ef[5], sf[6] = 60, 200
    rc, st, du, sc = oracle.align_run(m["tp"], m["sseq"], g["a_ssid"], g["a_tmatid"], scr, sf=sf, ef=ef)
    assert rc == 0 and du.sum() == len(scr) and not np.array_equal(st, g["a_start"])
