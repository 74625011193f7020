"""CPU: the Python mirror of the reference front end's initialisation (fe_tables.make_fe_desc)
against the tables inside the compiled reference's fe_t, bit for bit; frame counting."""
import os

import numpy as np
import pytest

from oracle import refdrv

pytestmark = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")
EN_US = os.path.join(os.path.dirname(refdrv.LIB_PATH), "model", "en-us")


def test_fe_tables_match_reference():
    from pocketsphinx_b200.fe_tables import make_fe_desc, n_frames
    m = refdrv.RefModel(EN_US)
    ref = m.fe_desc()
    ours = make_fe_desc()
    for k in ("frame_size", "frame_shift", "fft_size", "fft_order", "n_filt", "n_cep", "remove_dc", "remove_noise",
              "transform", "lifter_val", "window", "cmn"):
        assert ours[k] == ref[k], k
    for k in ("alpha", "sqrt_inv_n", "sqrt_inv_2n"):
        assert np.float32(ours[k]).tobytes() == np.float32(ref[k]).tobytes(), k
    for k in ("hamming", "ccc", "sss", "spec_start", "filt_start", "filt_width", "filt_coeffs", "mel_cosine", "lifter"):
        assert ours[k].dtype == ref[k].dtype and ours[k].shape == ref[k].shape, k
        assert ours[k].tobytes() == ref[k].tobytes(), k
    pcm = np.zeros(2000, np.int16)
    for n in (0, 1, 100, 409, 410, 411, 569, 570, 571, 2000):
        assert n_frames(ours, n) == len(m.mfcc(pcm[:n])), n
    m.close()
