"""Senone-dump (.sen) wire format (acmod.c:335-346, 880-1017): our writer/reader against the
reference's own writer and its ps_decode_senscr reader."""
import os

import numpy as np
import pytest

from conftest import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
have_ref = os.path.exists(os.path.join(REF, "libpsref.so")) and os.path.isdir(os.path.join(REF, "model", "en-us"))
ARGS = (os.path.join(REF, "model", "en-us"), os.path.join(REF, "model", "en-us.lm.bin"),
        os.path.join(REF, "model", "cmudict-en-us.dict"))


def test_roundtrip_and_partial_frames(tmp_path):
    from pocketsphinx_b200 import api
    rng = np.random.default_rng(0)
    scr = rng.integers(0, 700, (17, 333)).astype(np.int16)
    p = str(tmp_path / "a.sen")
    api.sendump_write(p, scr, mdef_file="/some/mdef")
    assert np.array_equal(api.sendump_read(p), scr)
    assert api.sendump_read(p, max_frames=5).shape == (5, 333)
    # a frame with a partial (delta-coded) list, as the reference writes without -compallsen
    with open(p, "ab") as f:
        ids = np.array([3, 4, 200], np.int64)
        f.write(np.int16(3).tobytes() + np.diff(np.concatenate([[0], ids])).astype(np.uint8).tobytes()
                + np.array([11, 12, 13], np.int16).tobytes())
    got = api.sendump_read(p)
    assert got.shape == (18, 333) and got[17, 3] == 11 and got[17, 4] == 12 and got[17, 200] == 13
    assert got[17, 5] == 0x7fff and got[17, 0] == 0x7fff


@pytest.mark.skipif(not have_ref, reason="oracle/_ref not present")
def test_writer_is_byte_identical_to_the_reference(tmp_path):
    from oracle import refdrv
    from pocketsphinx_b200 import api
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    ref_sen, our_sen = str(tmp_path / "ref.sen"), str(tmp_path / "ours.sen")
    a = refdrv.decode_senscr(*ARGS, pcm=pcm, senout=ref_sen, pl_window=0)
    assert a["hyp"] == "go forward ten meters"
    scr = api.sendump_read(ref_sen)
    assert np.array_equal(scr, golden("en_us_goforward.npz")["senscr"])
    api.sendump_write(our_sen, scr, mdef_file=os.path.join(ARGS[0], "mdef"))
    assert open(ref_sen, "rb").read() == open(our_sen, "rb").read()
    assert refdrv.decode_senscr(*ARGS, senfile=our_sen, pl_window=0)["hyp"] == "go forward ten meters"


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref, reason="oracle/_ref not present")
def test_gpu_scores_drive_the_reference_search(tmp_path, en_us):
    """GPU batch scores -> .sen -> the unmodified reference search (ps_decode_senscr)."""
    from oracle import refdrv
    from pocketsphinx_b200 import api
    g = golden("en_us_goforward.npz")
    m = api.Model(en_us)
    b = api.Batch(m, 2, 1024)
    scr = b.score_host(g["feats"], np.array([0, 278], np.int32))
    b.close(); m.close()
    ours, ref = str(tmp_path / "gpu.sen"), str(tmp_path / "gold.sen")
    api.sendump_write(ours, scr, mdef_file="mdef")
    api.sendump_write(ref, g["senscr"], mdef_file="mdef")
    assert open(ours, "rb").read() == open(ref, "rb").read()
    d = refdrv.decode_senscr(*ARGS, senfile=ours, pl_window=0)
    assert d["hyp"] == "go forward ten meters" and d["n_frames"] == 278
