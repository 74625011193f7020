"""CPU: the C restatement of allphone_search.c (oracle/ps_oracle.c: pso_allphone_run) plus the
backtrace against what the reference's own allphone_search produced on goforward.raw
(tests/golden/en_us_allphone.npz), on the reference's senone scores."""
import numpy as np

from conftest import golden


def test_allphone_oracle_matches_reference_golden():
    from oracle import oracle
    g, m = golden("en_us_allphone.npz"), golden("en_us_ptm_model.npz")
    scr = golden("en_us_goforward.npz")["senscr"]
    hist, n = oracle.allphone_run(m["tp"], m["sseq"], g["ssid"], g["tmatid"], g["succ_off"], g["succ"], int(g["start"]),
                                  int(g["beam"]), int(g["pbeam"]), int(g["inspen"]), scr)
    assert n == int(g["n_history"]) == len(hist)
    segs = oracle.allphone_backtrace(hist, g["ci"], len(scr) - 1, int(g["inspen"]))
    assert np.array_equal(segs, g["segs"])
    assert segs[0, 1] == 0 and segs[-1, 2] == len(scr) - 1 and (segs[1:, 1] == segs[:-1, 2] + 1).all()


def test_allphone_lm_oracle_matches_reference_golden():
    """With the shipped phone LM (dense score tables tabulated through the reference's LM object)."""
    from oracle import oracle
    g, m = golden("en_us_allphone.npz"), golden("en_us_ptm_model.npz")
    scr = golden("en_us_goforward.npz")["senscr"]
    hist, n = oracle.allphone_lm_run(m["tp"], m["sseq"], g["ssid"], g["tmatid"], g["succ_off"], g["succ"], int(g["start"]),
                                     int(g["beam"]), int(g["pbeam"]), g["ci"], g["lm_bg"], g["lm_tg"], scr)
    assert n == int(g["lm_n_history"])
    segs = oracle.allphone_backtrace_lm(hist, g["ci"], len(scr) - 1)
    assert np.array_equal(segs, g["lm_segs"])
    assert not np.array_equal(g["lm_segs"][:, :3], g["segs"][:len(g["lm_segs"]), :3])   # the LM changes the segmentation
