"""GPU (-m gpu): the batched device front end (row f-2) against the compiled reference's fe/ + feat/
run live on the same PCM (oracle/_ref/libpsref.so travels to the GPU box).  Everything up to the
log() is IEEE-reproducible and the tables are the reference's own; device log() vs glibc log() can
differ in the last bit of a float64, so parity is asserted at 1e-4 relative (north_star's bar for
float paths) and the share of bit-identical values is reported / bounded from below."""
import os

import numpy as np
import pytest

from oracle import refdrv

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")]
REF = os.path.dirname(refdrv.LIB_PATH)
EN_US = os.path.join(REF, "model", "en-us")


def _close(got, want, what):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if got.size == 0:
        return 1.0
    # features are differences of logs: compare against the scale of the cepstra, not of each value
    # (an utterance whose frames all have c0 < 0 gets a 0/0 CMN mean in the reference: NaN on both sides)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan), what
    if nan.all():
        return 1.0
    g, w = got[~nan].astype(np.float64), want[~nan].astype(np.float64)
    scale = max(1.0, float(np.abs(w).max()))
    err = np.abs(g - w).max()
    assert err <= 1e-4 * scale, "%s: max abs err %g (scale %g)" % (what, err, scale)
    return float((got.view(np.uint32) == want.view(np.uint32)).mean())


def _synth_pcm(rng, n):
    t = np.arange(n) / 16000.0
    x = rng.normal(0, 800, n)
    for _ in range(4):
        f0 = rng.uniform(100, 3500)
        a, b = sorted(rng.integers(0, max(n, 1), 2))
        x[a:b] += rng.uniform(1000, 6000) * np.sin(2 * np.pi * f0 * t[a:b])
    return np.clip(x, -32768, 32767).astype(np.int16)


@pytest.fixture(scope="module")
def api():
    from pocketsphinx_b200 import api
    assert api.device_count() > 0
    return api


def test_fe_goforward_and_ragged_batch_match_reference(api):
    from pocketsphinx_b200.fe_tables import make_fe_desc
    ref = refdrv.RefModel(EN_US)
    fe = api.FrontEnd(make_fe_desc())
    rng = np.random.default_rng(5)
    go = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    utts = [go, go[:0], go[:1], go[:100], go[:409], go[:410], go[:411], go[:570], go[:571], go[3000:20000]]
    utts += [_synth_pcm(rng, int(n)) for n in rng.integers(500, 50000, 12)]
    utts += [np.zeros(4000, np.int16), np.full(3000, -32768, np.int16)]          # silence (log floor), clipping
    off = api.FrontEnd.sample_offsets([len(u) for u in utts])
    feats, foff, mfcc = fe.process_host(np.concatenate(utts), off, want_mfcc=True)
    exact = []
    for u, pcm in enumerate(utts):
        want = ref.featurize_fresh(pcm) if len(pcm) else np.zeros((0, 39), np.float32)
        got = feats[foff[u]:foff[u + 1]]
        assert fe.n_frames(len(pcm)) == len(want) == len(got), "utterance %d (%d samples)" % (u, len(pcm))
        if len(pcm):
            exact.append(_close(got, want, "utterance %d (%d samples)" % (u, len(pcm))))
            assert np.array_equal(mfcc[foff[u]:foff[u + 1]], got[:, :13], equal_nan=True)
    assert feats.shape[0] == foff[-1]
    assert np.mean(exact) > 0.99, "share of bit-identical feature values %.4f" % np.mean(exact)
    # idempotent, batch composition does not matter
    f2, foff2 = fe.process_host(np.concatenate(utts[::-1]), api.FrontEnd.sample_offsets([len(u) for u in utts[::-1]]))
    for u in range(len(utts)):
        v = len(utts) - 1 - u
        assert np.array_equal(f2[foff2[v]:foff2[v + 1]], feats[foff[u]:foff[u + 1]], equal_nan=True)
    fe.close(); ref.close()


@pytest.mark.parametrize("kv,mk", [
    (dict(transform="legacy", remove_noise="no", lifter="0", nfilt="40", lowerf="133.33334", upperf="6855.4976"),
     dict(transform="legacy", remove_noise=False, lifter=0, nfilt=40, lowerf=133.33334, upperf=6855.4976)),
    (dict(transform="htk", remove_dc="yes", lifter="22"), dict(transform="htk", remove_dc=True, lifter=22)),
    (dict(cmn="none", remove_noise="yes", transform="dct"), dict(cmn="none")),
])
def test_fe_other_configurations(api, kv, mk):
    from pocketsphinx_b200.fe_tables import make_fe_desc
    ref = refdrv.RefModel(EN_US, **kv)
    desc = make_fe_desc(**mk)
    rd = ref.fe_desc()
    for k in ("remove_noise", "transform", "remove_dc", "lifter_val", "n_filt", "cmn"):
        assert desc[k] == rd[k], k
    fe = api.FrontEnd(rd)                                   # tables straight out of the reference's fe_t
    fe2 = api.FrontEnd(desc)                                # and from the Python mirror
    rng = np.random.default_rng(8)
    utts = [_synth_pcm(rng, int(n)) for n in (16000, 7777, 411, 30000)]
    off = api.FrontEnd.sample_offsets([len(u) for u in utts])
    feats, foff = fe.process_host(np.concatenate(utts), off)
    feats2, _ = fe2.process_host(np.concatenate(utts), off)
    assert np.array_equal(feats, feats2)
    for u, pcm in enumerate(utts):
        _close(feats[foff[u]:foff[u + 1]], ref.featurize_fresh(pcm), "utterance %d" % u)
    fe.close(); fe2.close(); ref.close()


def test_fe_feeds_the_scorer(api, en_us):
    """PCM -> device features -> device senone scores, against the reference scoring ITS features;
    identical features give identical int16 scores, so any mismatch is bounded by the feature check."""
    import torch
    from pocketsphinx_b200.fe_tables import make_fe_desc
    ref = refdrv.RefModel(EN_US)
    go = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    fe = api.FrontEnd(make_fe_desc())
    d_pcm = torch.from_numpy(go).cuda()
    T = fe.n_frames(len(go))
    d_feats = torch.empty((T, 39), dtype=torch.float32, device="cuda")
    foff, ms = fe.process_device(d_pcm.data_ptr(), np.array([0, len(go)], np.int64), d_feats.data_ptr())
    assert foff.tolist() == [0, 278] and ms > 0
    m = api.Model(en_us)
    b = api.Batch(m, 2, 512)
    b.score_device(d_feats.data_ptr(), foff)
    b.sync()
    want_feats = ref.featurize_fresh(go)
    got_feats = d_feats.cpu().numpy()
    scr = b.score_host(got_feats, foff)
    want = ref.score(got_feats)                              # reference GMM on OUR features: bit-exact
    assert np.array_equal(scr, want)
    same = (got_feats.view(np.uint32) == want_feats.view(np.uint32)).mean()
    assert same > 0.99
    b.close(); m.close(); fe.close(); ref.close()


def test_decode_from_pcm(api, en_us):
    """psb_decode_batch_pcm_host == front end, then psb_decode_batch_host on its features."""
    from pocketsphinx_b200.fe_tables import make_fe_desc
    go = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    utts = [go, go[5000:30000], go[:300]]
    off = api.FrontEnd.sample_offsets([len(u) for u in utts])
    fe = api.FrontEnd(make_fe_desc())
    m = api.Model(en_us)
    b = api.Batch(m, 8, 1024)
    ctx = api.HmmContext(en_us.tp, en_us.sseq, en_us.n_sen)
    H = en_us.n_ciphone
    pl = api.PhoneLoop(ctx, en_us.phone_ssid[:H], en_us.phone_tmat[:H], 5, -1080, -1080, 0, 3.0)
    foff, best, pen, scr = b.decode_pcm_host(fe, pl, np.concatenate(utts), off, want_senscr=True)
    feats, foff2 = fe.process_host(np.concatenate(utts), off)
    assert np.array_equal(foff, foff2)
    best2, pen2, scr2 = b.decode_host(pl, feats, foff2, want_senscr=True)
    assert np.array_equal(best, best2) and np.array_equal(pen, pen2) and np.array_equal(scr, scr2)
    b.close(); pl.close(); ctx.close(); m.close(); fe.close()
