"""CPU, build container only (skipped without oracle/_ref): synthetic models written as
Sphinx-3 files, loaded by the UNMODIFIED reference, must (1) come out as exactly the arrays our
loader mirrors produce and (2) score exactly like the C oracle -- this pins the oracle against the
reference on the BASELINE.json shapes that no shipped model covers (256 Gaussians / 5138 senones,
multi-Gaussian continuous with the wide log-add branch ms_senone.c:383-393, 8-bit semi)."""
import os

import numpy as np
import pytest

from oracle import oracle, refdrv
from pocketsphinx_b200 import s3io
from pocketsphinx_b200.model import PackedModel, synth_feats, synth_ms, synth_ptm, synth_semi

pytestmark = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")

FEAT_PTM = "-feat 1s_c_d_dd\n-svspec 0-12/13-25/26-38\n-cmn batch\n-agc none\n"
FEAT_SC = "-feat s2_4x\n-cmn batch\n-agc none\n"
FEAT_CONT = "-feat 1s_c_d_dd\n-cmn batch\n-agc none\n"


def _load(tmp_path, pm, raw, feat_params, **kv):
    d = str(tmp_path / "model")
    sen2ci = pm.sen2cb if pm.kind == "ptm" else np.concatenate(
        [np.repeat(np.arange(10), 3), np.arange(pm.n_sen - 30) % 10]).astype(np.int32)
    n_ci = pm.n_mgau if pm.kind == "ptm" else 10
    s3io.write_model_dir(d, kind=pm.kind, n_mgau=pm.n_mgau, n_feat=pm.n_feat, n_density=pm.n_density,
                         featlen=pm.featlen, mean=raw["mean"], var_raw=raw["var_raw"], tp_float=raw["tp_float"],
                         sen2ci=sen2ci, n_ci=n_ci, n_emit=3, n_ci_sen=n_ci * 3, mixw_q=raw.get("mixw_q"),
                         mixw_cb=raw.get("mixw_cb"), mixw_float=raw.get("mixw_float"), feat_params=feat_params)
    return refdrv.RefModel(d, **kv)


def _same_model(ref, pm):
    got = PackedModel.from_dict(ref.packed())
    assert got.kind == pm.kind
    for k in ("mean", "var", "det", "mixw", "mixw_cb", "sen2cb", "logadd8", "tp"):
        if k == "logadd8" and pm.kind == "ms":
            continue
        a, b = getattr(got, k), getattr(pm, k)
        assert a.shape == b.shape and np.array_equal(a, b), "model array %s differs after the reference loaded it" % k
    return got


def test_ptm_baseline_shape_through_reference(tmp_path):
    pm, raw = synth_ptm(seed=0, n_density=256, n_sen=5138, return_raw=True)
    ref = _load(tmp_path, pm, raw, FEAT_PTM)
    assert (ref.kind, ref.n_sen, ref.n_mgau, ref.n_density) == ("ptm", 5138, 42, 256)
    _same_model(ref, pm)
    feats = synth_feats(pm, 3, 40, seed=2)
    om = oracle.OracleModel(pm)
    for u in range(3):
        assert np.array_equal(ref.score(feats[u]), om.score_utt(feats[u]))
    ref.close()


@pytest.mark.parametrize("four_bit", [False, True])
def test_semi_synthetic_through_reference(tmp_path, four_bit):
    pm, raw = synth_semi(seed=1, n_sen=600, four_bit=four_bit, return_raw=True)
    ref = _load(tmp_path, pm, raw, FEAT_SC)
    assert ref.kind == "s2_semi" and ref.mixw_4bit == four_bit
    _same_model(ref, pm)
    feats = synth_feats(pm, 2, 30, seed=3)
    om = oracle.OracleModel(pm)
    for u in range(2):
        assert np.array_equal(ref.score(feats[u]), om.score_utt(feats[u]))
    ref.close()


@pytest.mark.parametrize("topn", [4, 2, 8])
def test_ms_multi_gaussian_continuous_through_reference(tmp_path, topn):
    pm, raw = synth_ms(seed=4, n_sen=400, n_density=8, topn=topn, return_raw=True)
    ref = _load(tmp_path, pm, raw, FEAT_CONT, senmgau=".cont.", topn=str(topn))
    assert ref.kind == "ms" and ref.n_mgau == 400 and ref.topn == topn
    got = _same_model(ref, pm)
    assert np.array_equal(got.logadd_ms, pm.logadd_ms) and got.logadd_ms_zero == pm.logadd_ms_zero
    feats = synth_feats(pm, 2, 25, seed=6)
    om = oracle.OracleModel(pm)
    for u in range(2):
        assert np.array_equal(ref.score(feats[u]), om.score_utt(feats[u]))
    ref.close()
