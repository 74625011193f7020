"""CPU (needs oracle/_ref/libpsref.so): the maintainer-side binding for the device searches
(integration/ps_search_cuda.c, compiled into the test copy of the reference) end to end:

  the reference decodes an utterance; its lextree / LM are flattened by cuda_fsg_export /
  cuda_ngram_export (every other grammar / n-gram test already goes through these); its OWN result tables
  are wiped; tables computed outside the reference -- by the device searches' phase code in host
  emulation here, by the kernels on a GPU -- are put back with cuda_fsg_import / cuda_ngram_import; and
  the reference's unchanged fsg_search_hyp / ngram_search_hyp (lattice construction + bestpath included)
  and segment iterator must give the hypothesis, score and word segmentation of an undisturbed decode."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import refdrv
from test_fsg_emul import _run as run_fsg
from test_fsg_emul import emul as fsg_emul  # noqa: F401
from test_ngf_emul import emuls, run_second  # noqa: F401
from test_ngs_emul import run_emul as run_first

pytestmark = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")
REF = os.path.dirname(refdrv.LIB_PATH)
HD = os.path.join(REF, "model", "en-us")
LM, DIC = os.path.join(REF, "data", "turtle.lm.bin"), os.path.join(REF, "data", "turtle.dic")


@pytest.fixture(scope="module")
def scored():
    ref = refdrv.RefModel(HD)
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    pk = ref.packed()
    scr = np.ascontiguousarray(ref.score(ref.featurize_fresh(pcm)))
    ref.close()
    ref = refdrv.RefModel(HD)                                   # a fresh object: the phone loop must see a fresh CMN state
    pl = ref.phoneloop(pcm)                                     # default look-ahead (window 5)
    ref.close()
    return pk, pcm, scr, pl


def test_grammar_tables_go_back_into_the_reference(fsg_emul, scored):  # noqa: F811
    pk, pcm, scr, _ = scored
    cmu = os.path.join(REF, "model", "cmudict-en-us.dict")
    fsg = os.path.join(ROOT, "tests", "golden", "commands.fsg")
    g = refdrv.fsg(HD, cmu, fsg, pcm)
    rows, n = run_fsg(fsg_emul, pk, g, scr, len(g["hist"]) + 16)
    assert n == len(g["hist"])
    rt = refdrv.fsg_roundtrip(HD, cmu, fsg, pcm, rows, len(scr))
    assert rt["n_entries"] == n and rt["hyp"] == g["hyp"] == "go forward ten meters" and rt["score"] == g["score"]
    # a table cut short gives a different (partial) answer: the import really is what the reference reads
    cut = int(np.searchsorted(rows[:, 1], 150))
    part = refdrv.fsg_roundtrip(HD, cmu, fsg, pcm, rows[:cut], 150)
    assert part["hyp"] != rt["hyp"]


@pytest.mark.parametrize("kv", [dict(fwdflat="yes", bestpath="yes", pl_window="5"),          # the shipped default pipeline
                                dict(fwdflat="yes", bestpath="no", pl_window="5"),
                                dict(fwdflat="no", bestpath="yes")])
def test_ngram_tables_go_back_into_the_reference(emuls, scored, kv):  # noqa: F811
    f1, f2 = emuls
    pk, pcm, scr, pl = scored
    want = refdrv.fwdtree(HD, LM, DIC, pcm, **kv)               # undisturbed decode, same configuration
    nc = want["n_ci"]
    la = dict(pl_pen=pl["pen"], pl_window=5) if "pl_window" in kv else {}
    n1, bp, bss, idx = run_first(f1, pk, want["info"], want["model"], scr, 8192, 1 << 18, **la)
    assert n1 > 0
    if kv["fwdflat"] == "yes":
        n2, bp, bss, idx = run_second(f2, pk, want["info"], want["model"], bp, scr, 8192, 1 << 18)
        assert n2 > 0
    assert np.array_equal(bp, want["bp"])                       # (what the other tests already establish)
    rt = refdrv.ngram_roundtrip(HD, LM, DIC, pcm, bp, bss, idx, **kv)
    assert rt["n_entries"] == len(bp)
    assert rt["hyp"] == want["hyp"] == "go forward ten meters" and rt["score"] == want["score"]
    segs = [ln.split() for ln in rt["seg"].splitlines()]
    assert [s[0] for s in segs if not s[0].startswith("<")] == want["hyp"].split()
    assert segs[0][1] == "0" and len(scr) - 3 <= int(segs[-1][2]) <= len(scr) - 1      # (without bestpath: the last frame that has exits)
    assert all(int(a[2]) + 1 == int(b[1]) for a, b in zip(segs, segs[1:]))
