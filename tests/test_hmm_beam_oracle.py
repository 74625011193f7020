"""oracle.sweep_beam (evaluate the active instances, prune to the beam, -maxhmmpf histogram) pinned on the compiled
reference: hmm_clear against the reference's own (hmm.c:181-196), and the whole loop run once over the oracle's
hmm_vit_eval and once over the reference's hmm_vit_eval + hmm_clear -- same records, same best scores, same counts.
The pruning rule itself restates ngram_search_fwdtree.c:1130-1181 / :811-827 / :872-874 (cited in the function)."""
import numpy as np
import pytest

from conftest import assert_hmm_equal, beam_case
from oracle import oracle, refdrv

pytestmark = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")


def test_hmm_clear_equals_the_reference():
    tp, sseq, hm, _ = beam_case(3, 64, 1)
    a, b = hm.copy(), hm.copy()
    for i in range(len(hm)):
        oracle.hmm_clear(a, i)
        refdrv.lib().refdrv_hmm_clear(b.ctypes.data, i)
    assert_hmm_equal(a, b, 3, "hmm_clear")


@pytest.mark.parametrize("n_emit,maxhmmpf", [(3, -1), (3, 150), (5, 90)])
def test_sweep_beam_oracle_equals_reference_functions(n_emit, maxhmmpf):
    tp, sseq, hm, n_sen = beam_case(n_emit, 400, 2 + n_emit)
    rng = np.random.default_rng(9)
    T = 12
    senscr = rng.integers(0, 900, (T, n_sen)).astype(np.int16)
    a, b = hm.copy(), hm.copy()
    octx, rctx = oracle.OracleHmmCtx(tp, sseq), refdrv.RefHmmCtx(tp, sseq)
    best_a, n_a = oracle.sweep_beam(octx, a, senscr, 7, -3000, maxhmmpf)
    best_b, n_b = oracle.sweep_beam(rctx, b, senscr, 7, -3000, maxhmmpf,
                                    clear=lambda h, i: refdrv.lib().refdrv_hmm_clear(h.ctypes.data, int(i)))
    rctx.close()
    assert np.array_equal(best_a, best_b) and np.array_equal(n_a, n_b)
    assert_hmm_equal(a, b, n_emit, "sweep_beam")
    # the pruning bites and is monotone; with -maxhmmpf the next frame never starts above the cap by more than one bin's ties
    assert n_a[0] > n_a[-1] > 0 and (np.diff(n_a) <= 0).all()
    inactive = hm["frame"] != 7
    assert_hmm_equal(a[inactive], hm[inactive], n_emit, "instances that were not active")
    assert ((a["frame"] == 7 + T) | (a["frame"] == -1) | inactive).all()
    if maxhmmpf >= 0:
        plain, _ = oracle.sweep_beam(octx, hm.copy(), senscr, 7, -3000, -1)
        n_plain = oracle.sweep_beam(octx, hm.copy(), senscr, 7, -3000, -1)[1]
        assert (n_a <= n_plain).all() and (n_a < n_plain).any()
