"""CPU, build container (needs oracle/_ref/libpsref.so): the oracle against the reference run LIVE
in configurations no committed fixture covers -- frame down-sampling (-ds), other phone-loop
windows / beams, -topn_beam for the semi-continuous back-end, tighter -pl_* settings."""
import os

import numpy as np
import pytest

from oracle import oracle, refdrv
from pocketsphinx_b200.model import PackedModel

pytestmark = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")
REF = os.path.dirname(refdrv.LIB_PATH)
GO = os.path.join(REF, "data", "goforward.raw")


@pytest.mark.parametrize("kv", [dict(ds="2"), dict(ds="3"), dict(ds="2", topn="4")])
def test_ptm_downsampling_matches_reference(kv):
    ref = refdrv.RefModel(os.path.join(REF, "model", "en-us"), **kv)
    pcm = np.fromfile(GO, np.int16)
    feats = ref.featurize(pcm)
    want, wtopn = ref.score(feats, want_topn=True)
    pm = PackedModel.from_dict(ref.packed())
    assert pm.ds_ratio == int(kv["ds"])
    got, topn = oracle.OracleModel(pm).score_utt(feats, want_topn=True)
    assert np.array_equal(topn, wtopn) and np.array_equal(got, want)
    ref.close()


@pytest.mark.parametrize("kv", [dict(pl_window="1"), dict(pl_window="2", pl_beam="1e-5", pl_pbeam="1e-3"),
                                dict(pl_window="9", pl_pip="0.5", pl_weight="1.5")])
def test_phoneloop_settings_match_reference(kv):
    ref = refdrv.RefModel(os.path.join(REF, "model", "en-us"))
    pcm = np.fromfile(GO, np.int16)
    pl = ref.phoneloop(pcm, **kv)
    pm = PackedModel.from_dict(ref.packed())
    par = pl["params"]
    H = par["n_phones"]
    o = oracle.phoneloop_run(pm.tp, pm.sseq, pm.phone_ssid[:H], pm.phone_tmat[:H], pl["senscr"], par["window"],
                             par["beam"], par["pbeam"], par["pip"], par["penalty_weight"])
    assert np.array_equal(o["best"], pl["best"])
    if par["window"] > 0:
        assert np.array_equal(o["pen"], pl["pen"])
    for k in ("score", "history", "out_score", "out_history", "bestscore", "frame"):
        assert np.array_equal(o["hmm"][k][..., :3] if o["hmm"][k].ndim == 3 else o["hmm"][k],
                              pl["hmm"][k][..., :3] if pl["hmm"][k].ndim == 3 else pl["hmm"][k]), k
    ref.close()


def test_semi_topn_beam_matches_reference():
    ref = refdrv.RefModel(os.path.join(REF, "model", "tidigits_hmm"), topn_beam="40")
    pcm = np.fromfile(GO, np.int16)
    feats = ref.featurize(pcm)
    want, wtopn = ref.score(feats, want_topn=True)
    pm = PackedModel.from_dict(ref.packed())
    assert pm.topn_beam.any()
    got, topn = oracle.OracleModel(pm).score_utt(feats, want_topn=True)
    assert np.array_equal(topn, wtopn) and np.array_equal(got, want)
    ref.close()
