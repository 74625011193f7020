"""CPU: the C restatement of kws_search.c (oracle/ps_oracle.c: pso_kws_run) plus the detection-list
logic against what the reference's own kws_search produced on goforward.raw
(tests/golden/en_us_kws.npz), on the reference's senone scores."""
import numpy as np
import pytest

from conftest import golden


@pytest.mark.parametrize("tag", ["a", "b"])
def test_kws_oracle_matches_reference_golden(tag):
    from oracle import oracle
    g, m = golden("en_us_kws.npz"), golden("en_us_ptm_model.npz")
    scr = golden("en_us_goforward.npz")["senscr"]
    hits = oracle.kws_run(m["tp"], m["sseq"], g[tag + "_pl_ssid"], g[tag + "_pl_tmat"], g[tag + "_kp_off"],
                          g[tag + "_kp_thresh"], g[tag + "_kp_ssid"], g[tag + "_kp_tmat"], int(g[tag + "_beam"]),
                          int(g[tag + "_plp"]), scr)
    assert len(hits) > len(g[tag + "_det"]) > 0                 # many raw hits collapse into few detections
    assert np.array_equal(oracle.kws_detections(hits), g[tag + "_det"])
