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


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref (compiled reference + models) not present")
@pytest.mark.parametrize("case", ["tidigits_semi", "an4_cont", "an4_cont_topn2"])
def test_full_decode_other_backends(case):
    """Same drop-in proof for the semi-continuous (tidigits, 4-bit sendump, s2_4x features) and the
    continuous ms back-end (an4_ci_cont), through the reference's own fwdtree/fwdflat search."""
    from oracle import refdrv
    from pocketsphinx_b200 import _lib
    M, D = os.path.join(REF, "model"), os.path.join(REF, "data")
    if case == "tidigits_semi":
        args = (os.path.join(M, "tidigits_hmm"), os.path.join(M, "tidigits_lm", "tidigits.lm.bin"),
                os.path.join(M, "tidigits_lm", "tidigits.dic"), np.fromfile(os.path.join(D, "dhd.2934z.raw"), np.int16))
        kv, backend = {}, "s2_semi"
    else:
        args = (os.path.join(M, "an4_ci_cont"), os.path.join(D, "turtle.lm.bin"), os.path.join(D, "turtle.dic"),
                np.fromfile(os.path.join(D, "goforward.raw"), np.int16))
        kv, backend = ({"topn": "2"} if case.endswith("topn2") else {}), "ms"
    cpu = refdrv.decode(*args, use_cuda=False, **kv)
    gpu = refdrv.decode(*args, use_cuda=True, libpath=_lib.LIB_PATH, **kv)
    assert cpu["hyp"] != "" and gpu["cuda_calls"] >= cpu["n_frames"]
    assert (gpu["hyp"], gpu["score"], gpu["seg"], gpu["n_frames"]) == (cpu["hyp"], cpu["score"], cpu["seg"], cpu["n_frames"])
