"""CPU: the plain-C oracle (oracle/ps_oracle.c) against fixtures produced by the compiled reference."""
import numpy as np
import pytest

from conftest import assert_hmm_equal, golden, hmm_view
from oracle import oracle


def test_ptm_goforward_bit_exact(en_us):
    g = golden("en_us_goforward.npz")
    om = oracle.OracleModel(en_us)
    scr, topn = om.score_utt(g["feats"], want_topn=True)
    assert scr.shape == (278, 5126)
    assert np.array_equal(scr, g["senscr"])
    assert np.array_equal(topn[0], g["topn_first"]) and np.array_equal(topn[-1], g["topn_last"])


def test_ptm_active_lists(en_us):
    g = golden("en_us_active.npz")
    gf = golden("en_us_goforward.npz")
    n_sen = int(g["n_sen"])
    flags = np.unpackbits(g["flags"], axis=1)[:, :n_sen]
    om = oracle.OracleModel(en_us)
    dec = om.decoder(n_hist=2)
    for t in range(flags.shape[0]):
        lst = oracle.flags2list(flags[t])
        assert len(lst) == g["nact"][t]
        scr = dec.frame_eval(gf["feats"][t], t, lst, compallsen=False)
        dec.set_frame_idx(t + 1)
        assert np.array_equal(scr, g["senscr"][t]), "frame %d" % t
    dec.close()


def test_flags2list_bridges_gaps():
    flags = np.zeros(2000, np.uint8)
    flags[[3, 4, 300, 1999]] = 1
    lst = oracle.flags2list(flags)
    # 3, +1, then 296 = 255 + 41, then 1699 = 6*255 + 169
    assert lst.tolist() == [3, 1, 255, 41] + [255] * 6 + [169]
    assert int(np.cumsum(lst.astype(np.int64))[-1]) == 1999


def test_semi_tidigits_bit_exact(tidigits):
    g = golden("tidigits_goforward.npz")
    assert tidigits.kind == "s2_semi" and tidigits.mixw_4bit
    om = oracle.OracleModel(tidigits)
    scr, topn = om.score_utt(g["feats"], want_topn=True)
    assert np.array_equal(topn, g["topn"])
    assert np.array_equal(scr, g["senscr"])


def test_ms_an4_bit_exact(an4):
    g = golden("an4_goforward.npz")
    assert an4.kind == "ms"
    om = oracle.OracleModel(an4)
    scr = om.score_utt(g["feats"])
    assert np.array_equal(scr, g["senscr"])


@pytest.mark.parametrize("n_emit", [3, 5, 4, 1])
def test_hmm_vit_eval(n_emit):
    g = golden("hmm_vit_eval.npz")
    ctx = oracle.OracleHmmCtx(g["n%d_tp" % n_emit], g["n%d_sseq" % n_emit])
    hm = hmm_view(g["n%d_before" % n_emit]).copy()
    want = hmm_view(g["n%d_after" % n_emit])
    best = ctx.vit_eval(hm, g["n%d_senscr" % n_emit])
    assert best == int(g["n%d_best" % n_emit])
    assert_hmm_equal(hm, want, n_emit, "n_emit=%d" % n_emit)


def test_phoneloop_goforward(en_us):
    g = golden("en_us_goforward.npz")
    n, beam, pbeam, pip, window = [int(x) for x in g["pl_params"]]
    o = oracle.phoneloop_run(en_us.tp, en_us.sseq, en_us.phone_ssid[:n], en_us.phone_tmat[:n], g["senscr"],
                             window, beam, pbeam, pip, float(g["pl_weight"]))
    assert np.array_equal(o["best"], g["pl_best"])
    assert np.array_equal(o["pen"], g["pl_pen"])
    assert_hmm_equal(o["hmm"], hmm_view(g["pl_hmm"]), 3, "phone loop")
