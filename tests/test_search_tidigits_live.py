"""CPU (needs oracle/_ref/libpsref.so and its copy of the tidigits model): all three searches on the
reference's OTHER shipped acoustic model -- semi-continuous, 5-state HMMs, its own dictionary,
grammar (tidigits.fsg) and LM (tidigits.lm.bin) -- run live: the reference against the oracle
restatements and against the device searches' phase code in both thread orders.  Covers the 5-state
topology and a 34-phone set in fsg_search, ngram_search_fwdtree and ngram_search_fwdflat."""
import os

import numpy as np
import pytest

from oracle import oracle, refdrv
from test_fsg_emul import _run as run_fsg
from test_fsg_emul import emul as fsg_emul  # noqa: F401
from test_ngf_emul import emuls, run_second  # noqa: F401
from test_ngs_emul import run_emul as run_first

pytestmark = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")
REF = os.path.dirname(refdrv.LIB_PATH)
HD = os.path.join(REF, "model", "tidigits_hmm")
DIC = os.path.join(REF, "model", "tidigits_lm", "tidigits.dic")
FSG = os.path.join(REF, "model", "tidigits_lm", "tidigits.fsg")
LM = os.path.join(REF, "model", "tidigits_lm", "tidigits.lm.bin")


@pytest.fixture(scope="module")
def scored():
    ref = refdrv.RefModel(HD)
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    pk = ref.packed()
    scr = np.ascontiguousarray(ref.score(ref.featurize_fresh(pcm)))
    ref.close()
    assert pk["tp"].shape[1] == 5                                # 5 emitting states
    return pk, pcm, scr


def test_grammar_search_on_tidigits(fsg_emul, scored):  # noqa: F811
    pk, pcm, scr = scored
    r = refdrv.fsg(HD, DIC, FSG, pcm)
    assert len(r["hist"]) > 5000
    got = oracle.fsg_run(pk["tp"], pk["sseq"], r, scr)
    assert got.shape == r["hist"].shape and np.array_equal(got, r["hist"])
    hist, n = run_fsg(fsg_emul, pk, r, scr, len(r["hist"]) + 16)
    assert n == len(r["hist"]) and np.array_equal(hist, r["hist"])
    bp, score = oracle.fsg_find_exit(got, r["links"], len(scr), r["final_state"])
    assert score == r["score"]


@pytest.mark.parametrize("kv", [dict(), dict(pl_window="5"), dict(beam="1e-60", wbeam="1e-40", maxwpf="5")])
def test_both_ngram_passes_on_tidigits(emuls, scored, kv):  # noqa: F811
    f1, f2 = emuls
    pk, pcm, scr = scored
    first = refdrv.fwdtree(HD, LM, DIC, pcm, **kv)
    both = refdrv.fwdtree(HD, LM, DIC, pcm, fwdflat="yes", **kv)
    nc = both["n_ci"]
    la = {}
    if "pl_window" in kv:
        ref = refdrv.RefModel(HD)
        pl = ref.phoneloop(pcm, pl_window=kv["pl_window"])
        ref.close()
        la = dict(pl_pen=pl["pen"], pl_window=int(kv["pl_window"]))
    o1 = oracle.fwdtree_run(pk["tp"], pk["sseq"], pk["phone_tmat"][:nc], both["info"], both["model"], scr, **la)
    assert np.array_equal(o1[0], first["bp"]) and np.array_equal(o1[1], first["bss"]) and np.array_equal(o1[2], first["bp_idx"])
    o2 = oracle.fwdflat_run(pk["tp"], pk["sseq"], pk["phone_tmat"][:nc], pk["phone_ssid"][:nc], both["info"], both["model"], o1[0], scr)
    assert np.array_equal(o2[0], both["bp"]) and np.array_equal(o2[1], both["bss"]) and np.array_equal(o2[2], both["bp_idx"])
    n1, bp1, bss1, idx1 = run_first(f1, pk, both["info"], both["model"], scr, len(first["bp"]) + 8, len(first["bss"]) + 64, **la)
    assert n1 == len(first["bp"]) and np.array_equal(bp1, first["bp"]) and np.array_equal(bss1, first["bss"])
    n2, bp2, bss2, idx2 = run_second(f2, pk, both["info"], both["model"], bp1, scr, len(both["bp"]) + 8, len(both["bss"]) + 64)
    assert n2 == len(both["bp"]) and np.array_equal(bp2, both["bp"]) and np.array_equal(bss2, both["bss"])
    assert np.array_equal(idx2, both["bp_idx"])
