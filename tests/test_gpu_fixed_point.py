"""GPU (-m gpu): the FIXED_POINT build's arithmetic (VERDICT row X1) through the C ABI:
psb_model_desc_t.fixed_point = 1 switches the Gaussian stage to Q12 integers -- FIXMUL
(fe/fixpoint.h:98-100), GMMSUB as gcc compiles it (tied_mgau_common.h:62-66), the observable early exits
of eval_cb (ptm_mgau.c:182-206, s2_semi_mgau.c:137-143).  Expected values: the reference compiled with
-DFIXED_POINT on goforward.raw (tests/golden/fx_*.npz, oracle/make_golden.py:make_fixed_point), the C
restatement (pinned on the same in tests/test_fixed_point_oracle.py) on synthetic wrap-stress data, and a
full decode of the FIXED_POINT reference with its back-end swapped.  All byte-identical."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, fx_case

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref")
FX = os.path.join(REF, "libpsref_fx.so")


@pytest.fixture(scope="module")
def api():
    from pocketsphinx_b200 import api
    assert api.device_count() > 0, "no CUDA device visible"
    return api


@pytest.mark.parametrize("name", ["en_us", "tidigits"])
def test_fx_batch_matches_fixed_point_reference(api, name):
    pm, feats, want, want_topn = fx_case(name)
    m = api.Model(pm)
    b = api.Batch(m, 8, 2048)
    T = len(feats)
    scr = b.score_host(feats, np.array([0, T], np.int32))
    assert np.array_equal(scr, want), "first differing frame %d" % int(np.argwhere((scr != want).any(1))[0, 0])
    # ragged batch: prefixes of the utterance restart from the post-init lists
    lens = [T, 0, 1, 17, 100]
    off = api.Batch.offsets(lens)
    scr = b.score_host(np.concatenate([feats[:n] for n in lens]), off)
    for u, n in enumerate(lens):
        assert np.array_equal(scr[off[u]:off[u + 1]], want[:n]), "utterance %d" % u
    b.close()
    m.close()


@pytest.mark.parametrize("name", ["en_us", "tidigits"])
def test_fx_scorer_frame_eval(api, name):
    """The per-frame drop-in (ps_mgaufuncs_t.frame_eval) in fixed-point arithmetic, history re-scoring included."""
    pm, feats, want, _ = fx_case(name)
    m = api.Model(pm)
    s = api.Mgau(m, pl_window=0)
    for t in range(80):
        scr = s.frame_eval(feats[t], t)
        s.frame_idx = t + 1
        assert np.array_equal(scr, want[t]), "frame %d" % t
    assert np.array_equal(s.frame_eval(feats[79], 79), want[79])
    s.close()
    m.close()


def _fx_synth(kind, seed, scale):
    """A synthetic model converted like the FIXED_POINT loaders do (means FLOAT2MFCC = x * 4096 truncated,
    variance terms and determinants truncated to int32), features in Q12; `scale` > 1 drives the squared
    differences past 2^31 so that FIXMUL truncates and GMMSUB wraps / floors."""
    from pocketsphinx_b200.model import synth_feats, synth_ptm, synth_semi
    pm = synth_ptm(seed=seed, n_density=128, n_sen=700) if kind == "ptm" else synth_semi(seed=seed, four_bit=(seed & 1) == 1)
    feats = synth_feats(pm, 6, 40, seed=seed + 1) * np.float32(scale)
    pm.mean = (pm.mean * np.float32(4096)).astype(np.int32).view(np.float32)
    pm.var = pm.var.astype(np.int32).view(np.float32)
    pm.det = pm.det.astype(np.int32).view(np.float32)
    pm.fixed_point = 1
    q = np.clip(feats.astype(np.float64) * 4096, -2**31, 2**31 - 1).astype(np.int32)
    return pm, q


@pytest.mark.parametrize("kind,seed,scale", [("ptm", 3, 1.0), ("ptm", 4, 40.0), ("ptm", 5, 3000.0),
                                             ("semi", 6, 1.0), ("semi", 7, 60.0), ("semi", 8, 3000.0)])
def test_fx_synthetic_matches_oracle(api, kind, seed, scale):
    from oracle import oracle
    pm, q = _fx_synth(kind, seed, scale)
    om = oracle.OracleModel(pm)
    m = api.Model(pm)
    b = api.Batch(m, 8, 1024)
    U, T = q.shape[:2]
    off = api.Batch.offsets([T] * U)
    scr = b.score_host(q.reshape(U * T, -1).view(np.float32), off)
    s = api.Mgau(m, pl_window=0)
    for u in range(U):
        want = om.score_utt(q[u].view(np.float32))
        assert np.array_equal(scr[off[u]:off[u + 1]], want), "utterance %d" % u
        if u == 0:
            for t in range(T):
                got = s.frame_eval(q[0, t].view(np.float32), t)
                s.frame_idx = t + 1
                assert np.array_equal(got, want[t]), "scorer frame %d" % t
    s.close(); b.close(); m.close()


def test_fx_model_rejects_ms(api):
    from pocketsphinx_b200.model import synth_ms
    pm = synth_ms(seed=1, n_sen=64)
    pm.fixed_point = 1
    with pytest.raises(Exception, match="fixed-point"):
        api.Model(pm)


DECODE = r"""
import sys, json, numpy as np
sys.path.insert(0, %r)
from oracle import refdrv
from pocketsphinx_b200 import _lib
args = (%r, %r, %r, np.fromfile(%r, np.int16))
cpu = refdrv.decode(*args, use_cuda=False)
gpu = refdrv.decode(*args, use_cuda=True, libpath=_lib.LIB_PATH)
print(json.dumps({"cpu": [cpu["hyp"], cpu["score"], cpu["seg"], cpu["n_frames"]],
                  "gpu": [gpu["hyp"], gpu["score"], gpu["seg"], gpu["n_frames"]], "calls": gpu["cuda_calls"]}))
"""


@pytest.mark.skipif(not os.path.exists(FX), reason="oracle/_ref/libpsref_fx.so not built (make -C oracle fx)")
@pytest.mark.parametrize("case", ["en_us_ptm", "tidigits_semi"])
def test_fx_full_decode_identical_with_cuda_backend(case):
    """The FIXED_POINT reference decodes with its own back-end and with the CUDA one bound through
    integration/ps_mgau_cuda.c (which no longer refuses the build): hypothesis, score, segments."""
    import json
    M, D = os.path.join(REF, "model"), os.path.join(REF, "data")
    if case == "en_us_ptm":
        a = (os.path.join(M, "en-us"), os.path.join(M, "en-us.lm.bin"), os.path.join(M, "cmudict-en-us.dict"),
             os.path.join(D, "goforward.raw"))
    else:
        a = (os.path.join(M, "tidigits_hmm"), os.path.join(M, "tidigits_lm", "tidigits.lm.bin"),
             os.path.join(M, "tidigits_lm", "tidigits.dic"), os.path.join(D, "dhd.2934z.raw"))
    r = subprocess.run([sys.executable, "-c", DECODE % ((ROOT,) + a)], env=dict(os.environ, PSREF_LIB=FX),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["cpu"][0] != "" and out["calls"] >= out["cpu"][3], "the CUDA back-end did not serve the decode"
    assert out["gpu"] == out["cpu"]
