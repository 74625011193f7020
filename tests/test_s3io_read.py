"""CPU: pocketsphinx_b200.s3io.read_model_dir -- an acoustic-model directory read WITHOUT the reference
(binary / text mdef, means, variances, transition matrices, sendump / mixture weights, feat.params) must
give, array for array, what the compiled reference holds after acmod_init on the same directory: its three
shipped models (PTM en-us, semi-continuous tidigits with 5-state HMMs, continuous an4) live, synthetic
4-bit / 8-bit sendumps through write -> read, and damaged files as errors."""
import os
import struct

import numpy as np
import pytest

from oracle import refdrv
from pocketsphinx_b200 import s3io
from pocketsphinx_b200.model import PackedModel, synth_ms, synth_semi

live = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")
REF = os.path.dirname(refdrv.LIB_PATH)
# bookkeeping fields the reference driver reports as 0 / empty for back-ends that do not have them
NOT_APPLICABLE = {"ptm": {"aw"}, "s2_semi": {"aw"}, "ms": {"ds_ratio", "logadd8"}}


@live
@pytest.mark.parametrize("name,kind", [("en-us", "ptm"), ("tidigits_hmm", "s2_semi"), ("an4_ci_cont", "ms")])
def test_shipped_models_read_like_the_reference_loads_them(name, kind):
    hd = os.path.join(REF, "model", name)
    got = s3io.read_model_dir(hd)
    ref = refdrv.RefModel(hd)
    want = ref.packed()
    ref.close()
    assert got["kind"] == want["kind"] == kind
    for k, v in want.items():
        if k in NOT_APPLICABLE[kind]:
            continue
        g = got[k]
        if isinstance(v, np.ndarray):
            g = np.asarray(g)
            assert g.size == v.size and np.array_equal(g.ravel(), v.ravel()), k
        else:
            assert g == v, k
    pm = PackedModel.from_dict(got)                              # and it is a model the scorers take
    assert pm.n_sen == want["n_sen"] and pm.sumlen == int(np.sum(want["featlen"]))


@live
def test_configuration_overrides_reach_the_loaders_arithmetic():
    hd = os.path.join(REF, "model", "en-us")
    kv = dict(varfloor="0.01", tmatfloor="0.3", topn="2", ds="2")
    got = s3io.read_model_dir(hd, **kv)
    ref = refdrv.RefModel(hd, **kv)
    want = ref.packed()
    ref.close()
    base = s3io.read_model_dir(hd)
    for k in ("var", "det", "tp"):
        assert np.array_equal(np.asarray(got[k]).ravel(), want[k].ravel()), k
        assert not np.array_equal(np.asarray(got[k]).ravel(), np.asarray(base[k]).ravel()), k
    assert (got["topn"], got["ds_ratio"]) == (want["topn"], want["ds_ratio"]) == (2, 2)


def _write(tmp_path, pm, raw, n_ci):
    d = str(tmp_path / "model")
    n_sen = pm.n_sen
    assert n_sen == n_ci * 3                                    # CI-only definition: a text mdef without triphones
    s3io.write_model_dir(d, kind=pm.kind, n_mgau=pm.n_mgau, n_feat=pm.n_feat, n_density=pm.n_density, featlen=pm.featlen,
                         mean=raw["mean"], var_raw=raw["var_raw"], tp_float=raw["tp_float"],
                         sen2ci=np.repeat(np.arange(n_ci), 3).astype(np.int32), n_ci=n_ci, n_emit=3, n_ci_sen=n_sen,
                         mixw_q=raw.get("mixw_q"), mixw_cb=raw.get("mixw_cb"), mixw_float=raw.get("mixw_float"), feat_params="-topn 4\n")
    return d


def _same(got, pm, keys=("mean", "var", "det", "mixw", "mixw_cb", "sen2cb", "tp")):
    got = PackedModel.from_dict(got)
    assert got.kind == pm.kind
    for k in keys:
        a, b = getattr(got, k), getattr(pm, k)
        assert a.shape == b.shape and np.array_equal(a, b), k


@pytest.mark.parametrize("four_bit", [False, True])
def test_semi_sendump_write_read_round_trip(tmp_path, four_bit):
    pm, raw = synth_semi(seed=3, n_sen=30, four_bit=four_bit, return_raw=True)
    got = s3io.read_model_dir(_write(tmp_path, pm, raw, 10))
    _same(got, pm)
    assert (len(got["mixw_cb"]) == 16) == four_bit and np.array_equal(got["topn_beam"], np.zeros(pm.n_feat, np.uint8))
    assert np.array_equal(got["sseq"], np.arange(30, dtype=np.uint16).reshape(10, 3))


def test_continuous_write_read_round_trip(tmp_path):
    pm, raw = synth_ms(seed=5, n_sen=30, n_density=4, return_raw=True)
    got = s3io.read_model_dir(_write(tmp_path, pm, raw, 10), topn=9)
    _same(got, pm, keys=("mean", "var", "det", "mixw", "sen2cb", "tp"))
    assert got["topn"] == 4 and np.array_equal(got["logadd_ms"], pm.logadd_ms)   # topn clamps to the densities (ms_mgau.c:144)


def test_damaged_files_are_errors(tmp_path):
    pm, raw = synth_semi(seed=6, n_sen=30, return_raw=True)
    d = _write(tmp_path, pm, raw, 10)
    s3io.read_model_dir(d)

    def damaged(name, fn):
        p = os.path.join(d, name)
        keep = open(p, "rb").read()
        open(p, "wb").write(fn(bytearray(keep)))
        try:
            with pytest.raises((ValueError, struct.error, NotImplementedError)):
                s3io.read_model_dir(d)
        finally:
            open(p, "wb").write(keep)

    def flip(b):
        b[len(b) // 2] ^= 0x40
        return bytes(b)
    damaged("means", flip)                                       # checksum
    damaged("transition_matrices", flip)
    damaged("variances", lambda b: bytes(b[:len(b) // 2]))
    damaged("sendump", lambda b: bytes(b[:len(b) - 100]))        # rows missing
    damaged("means", lambda b: b"s4\n" + bytes(b[3:]))
    damaged("mdef", lambda b: bytes(b).replace(b"0 n_tri", b"3 n_tri"))   # triphones in a text mdef: say so, do not guess
    damaged("mdef", lambda b: bytes(b).replace(b"0.3", b"0.9", 1))
    s3io.read_model_dir(d)


@pytest.mark.parametrize("name,gold", [("en-us", "en_us_ptm_model.npz"), ("tidigits_hmm", "tidigits_sc_model.npz"),
                                       ("an4_ci_cont", "an4_cont_model.npz")])
def test_directory_read_equals_the_golden_model_the_gpu_tests_score_with(name, gold):
    """The GPU parity tests run on model arrays exported from the compiled reference (tests/golden/*_model.npz);
    a directory read here gives the same arrays, so everything verified on those holds for PackedModel.from_dir."""
    from conftest import GOLDEN
    hd = os.path.join(REF, "model", name)
    if not os.path.isdir(hd) or not os.path.exists(os.path.join(GOLDEN, gold)):
        pytest.skip("model directory or golden file not present")
    pm = PackedModel.from_dir(hd)
    want = PackedModel.load(os.path.join(GOLDEN, gold))
    assert pm.kind == want.kind
    for k in ("featlen", "mean", "var", "det", "mixw", "mixw_cb", "sen2cb", "tp", "sseq", "phone_ssid", "phone_tmat", "topn_beam"):
        a, b = getattr(pm, k), getattr(want, k)
        if k.startswith("phone_"):
            a = a[:len(b)]                                       # the golden file keeps the first few thousand phones
        assert a.shape == b.shape and np.array_equal(a, b), k
    assert (pm.n_sen, pm.n_mgau, pm.n_density, pm.topn, pm.n_emit_state) == (want.n_sen, want.n_mgau, want.n_density, want.topn, want.n_emit_state)


def test_other_endian_files_read_the_same(tmp_path):
    """bio_readhdr's byte-order word (bio.c:232-262): files written on a big-endian machine."""
    pm, raw = synth_ms(seed=7, n_sen=30, n_density=4, return_raw=True)
    d = _write(tmp_path, pm, raw, 10)
    want = s3io.read_model_dir(d)
    for name in ("means", "variances", "transition_matrices", "mixture_weights"):
        p = os.path.join(d, name)
        b = open(p, "rb").read()
        cut = b.index(b"endhdr\n") + 7
        body = np.frombuffer(b[cut:], "<u4").astype(">u4").tobytes()      # every 32-bit word, magic and checksum included
        open(p, "wb").write(b[:cut] + body)
    got = s3io.read_model_dir(d)
    for k in ("mean", "var", "det", "mixw", "tp"):
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), k
