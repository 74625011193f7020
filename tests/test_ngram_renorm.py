"""CPU: score renormalisation of the n-gram search (renormalize_scores ngram_search_fwdtree.c:566-602,
fwdflat_renormalize_scores ngram_search_fwdflat.c:785-810) in the device kernels' phase code.

The reference's condition, best_score + 2 * beam < WORST_SCORE, is a window of 2 |beam| above the int32
floor that the best score has to land in.  With >> 10 scores that takes hundreds of thousands of frames
(or a search that has long since lost its discriminating power), so it cannot be provoked through the
reference's own configuration in a test-sized run: -beam is a probability and its logarithm bottoms out
near -7 000, far from the floor at -536 870 912.  What CAN be checked: the phase code against the
oracle's literal restatement of the two functions on a search description whose beam field is set below
WORST_SCORE / 2 by hand, which makes the condition true in EVERY frame (renormalise, evaluate, prune with
a beam that prunes nothing).  Both thread orders reproduce the oracle's tables row for row; the oracle's
renormalisation branch itself is the five lines of hmm_normalize calls of the reference, unpinned by a
live run for the reason above."""
import numpy as np

from conftest import golden
from test_ngs_emul import _case, emul, run_emul  # noqa: F401  (the emulation fixture)


def test_first_pass_phase_code_renormalises_like_the_oracle(emul):
    from oracle import oracle
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    c = _case(golden("en_us_fwdtree.npz"), "default")
    info = c["info"].copy()
    info[8] = -300000000                    # beam: best_score + 2 * beam < WORST_SCORE in every frame
    info[14] = 3                            # -maxwpf keeps the tables small although nothing is pruned by the beam
    scr = gf["senscr"][:60]
    want = oracle.fwdtree_run(m["tp"], m["sseq"], m["phone_tmat"][:int(info[6])], info, c["model"], scr)
    base = oracle.fwdtree_run(m["tp"], m["sseq"], m["phone_tmat"][:int(c["info"][6])], c["info"], c["model"], scr)
    assert len(want[0]) > 0 and int(np.abs(want[0][:, 4]).max()) < int(np.abs(base[0][:, 4]).max())   # scores were pulled back to 0
    n, bp, bss, idx = run_emul(emul, m, info, c["model"], scr, len(want[0]) + 8, len(want[1]) + 64)
    assert n == len(want[0]) and np.array_equal(bp, want[0])
    assert np.array_equal(bss, want[1]) and np.array_equal(idx, want[2])


def test_second_pass_phase_code_renormalises_like_the_oracle(tmp_path_factory):
    import test_ngf_emul as ngf
    from oracle import oracle
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    c = ngf._case(golden("en_us_fwdtree.npz"), "flat_wide")
    nci = int(c["info"][6])
    scr = gf["senscr"][:90]
    bp1 = oracle.fwdtree_run(m["tp"], m["sseq"], m["phone_tmat"][:nci], c["info"], c["model"], scr)[0]     # a normal first pass
    info = c["info"].copy()
    info[8] = -300000000                    # the second pass tests the same field (ngram_search_fwdflat.c:830)
    want = oracle.fwdflat_run(m["tp"], m["sseq"], m["phone_tmat"][:nci], m["phone_ssid"][:nci], info, c["model"], bp1, scr)
    base = oracle.fwdflat_run(m["tp"], m["sseq"], m["phone_tmat"][:nci], m["phone_ssid"][:nci], c["info"], c["model"], bp1, scr)
    assert len(want[0]) > 0 and int(np.abs(want[0][:, 4]).max()) < int(np.abs(base[0][:, 4]).max())
    for reverse in (False, True):
        f2 = ngf._build(tmp_path_factory.mktemp("ngfrenorm"), "ngf_%d" % reverse, "ngf_emul.cpp", reverse).ngf_emul_run
        f2.restype = ngf.C.c_int32
        f2.argtypes = ngf.ARGT
        n, bp, bss, idx = ngf.run_second(f2, m, info, c["model"], bp1, scr, len(want[0]) + 8, len(want[1]) + 64)
        assert n == len(want[0]) and np.array_equal(bp, want[0])
        assert np.array_equal(bss, want[1]) and np.array_equal(idx, want[2])
