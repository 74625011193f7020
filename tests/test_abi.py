"""CPU: the C-ABI library builds, loads, and exports every symbol include/psb200.h declares."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pocketsphinx_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "psb200.h")).read()
    declared = set(re.findall(r"\b(psb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"psb_status_e"}
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(L, name)
    assert L.psb_abi_version() == 1


def test_hmm_struct_is_88_bytes():
    from pocketsphinx_b200.api import HMM_DTYPE
    assert HMM_DTYPE.itemsize == 88
    assert HMM_DTYPE.fields["bestscore"][1] == 68 and HMM_DTYPE.fields["frame"][1] == 76


def test_no_cpu_fallback_without_gpu():
    """Without a device the product must fail loudly, not compute on the CPU."""
    import pytest
    from pocketsphinx_b200 import api
    from pocketsphinx_b200.model import synth_ptm
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.PsbError):
        api.Model(synth_ptm(n_density=32, n_sen=200))
