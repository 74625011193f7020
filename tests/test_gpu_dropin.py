"""GPU (-m gpu): the drop-in proof.  The UNMODIFIED reference decoder (compiled into
oracle/_ref/libpsref.so) decodes goforward.raw twice -- once with its own ptm back-end, once with
acmod->mgau replaced by the CUDA back-end bound through integration/ps_mgau_cuda.c -- and the
hypothesis, its score and every word segment must be identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _have_ref():
    return os.path.exists(os.path.join(REF, "libpsref.so")) and os.path.isdir(os.path.join(REF, "model", "en-us"))


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref (compiled reference + en-us model) not present")
@pytest.mark.parametrize("kv", [{}, {"fwdflat": "no", "bestpath": "no"}, {"pl_window": "0"}, {"ds": "2"}])
def test_full_decode_identical_with_cuda_backend(kv):
    from oracle import refdrv
    from pocketsphinx_b200 import _lib
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    args = (os.path.join(REF, "model", "en-us"), os.path.join(REF, "model", "en-us.lm.bin"),
            os.path.join(REF, "model", "cmudict-en-us.dict"), pcm)
    cpu = refdrv.decode(*args, use_cuda=False, **kv)
    gpu = refdrv.decode(*args, use_cuda=True, libpath=_lib.LIB_PATH, **kv)
    assert cpu["hyp"] == "go forward ten meters"
    assert gpu["cuda_calls"] >= cpu["n_frames"], "the CUDA back-end did not serve the decode"
    assert gpu["hyp"] == cpu["hyp"]
    assert gpu["score"] == cpu["score"]
    assert gpu["seg"] == cpu["seg"]
    assert gpu["n_frames"] == cpu["n_frames"]
