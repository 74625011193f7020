"""GPU: pocketsphinx_b200.decoder.Decoder --
audio in, words out, everything read from the reference's files by the package itself -- against the
hypothesis, score and segmentation the reference produces for the same configuration (golden: the tables of
tests/golden/en_us_fwdtree.npz flat_default; live when oracle/_ref/libpsref.so is there)."""
import os

import numpy as np
import pytest

from conftest import ROOT, golden

pytestmark = [pytest.mark.gpu]
REF = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.timeout(600)
def test_decoder_audio_to_words():
    from pocketsphinx_b200 import api
    from pocketsphinx_b200.decoder import Decoder
    hd, dic, lm = os.path.join(REF, "model", "en-us"), os.path.join(REF, "data", "turtle.dic"), os.path.join(REF, "data", "turtle.lm.bin")
    if not os.path.exists(lm):
        pytest.skip("reference data files not present")
    go = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    dec = Decoder(hd, dic, lm, max_utts=8, max_frames=4096)
    out = dec.decode_raw_batch([go, go[:30000], go])
    assert out[0]["hyp"] == out[2]["hyp"] == "go forward ten meters"
    assert out[0]["score"] == out[2]["score"] and np.array_equal(out[0]["seg"], out[2]["seg"])
    # the device front end agrees with the reference's to 1e-4 relative, not bit for bit, so the golden tables (computed
    # on the reference's own features) pin words and segment boundaries, not every score
    g = golden("en_us_fwdtree.npz")
    want = {k[len("flat_default."):]: g[k] for k in g.files if k.startswith("flat_default.")}
    e, s, chain = api.ngram_hyp(want["bp"], want["bp_idx"], len(want["bp_idx"]) - 1, int(want["info"][20]))
    assert [int(w) for w in out[0]["seg"][:, 1]] == [int(w) for w in chain[:, 1]]
    assert np.abs(out[0]["seg"][:, 3] - chain[:, 3]).max() <= 2 and abs(out[0]["score"] - s) < 200
    assert out[1]["n_frames"] < out[0]["n_frames"] and out[1]["hyp"] != ""
    dec.close()
