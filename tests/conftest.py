import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def en_us():
    from pocketsphinx_b200.model import PackedModel
    return PackedModel.load(os.path.join(GOLDEN, "en_us_ptm_model.npz"))


@pytest.fixture(scope="session")
def tidigits():
    from pocketsphinx_b200.model import PackedModel
    return PackedModel.load(os.path.join(GOLDEN, "tidigits_sc_model.npz"))


@pytest.fixture(scope="session")
def an4():
    from pocketsphinx_b200.model import PackedModel
    return PackedModel.load(os.path.join(GOLDEN, "an4_cont_model.npz"))


def hmm_view(raw):
    """uint8 [..., 88] golden dump -> structured hmm_t array."""
    from oracle.oracle import HMM_DTYPE
    return np.ascontiguousarray(raw).view(HMM_DTYPE).reshape(raw.shape[:-1])


HMM_FIELDS = ["score", "history", "out_score", "out_history", "ssid", "senid", "bestscore", "tmatid",
              "frame", "mpx", "n_emit_state"]


def assert_hmm_equal(a, b, n_emit, what=""):
    for k in HMM_FIELDS:
        x, y = a[k], b[k]
        if k in ("score", "history", "senid"):
            x, y = x[..., :n_emit], y[..., :n_emit]
        assert np.array_equal(x, y), "%s: hmm field %s differs at %s" % (
            what, k, np.argwhere(x != y)[:5].tolist())
