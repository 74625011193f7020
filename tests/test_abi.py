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


def test_no_fused_multiply_add_in_distance_kernels():
    """The Gaussian exponent must round the product and the subtraction separately
    (ptm_mgau.c:64-69).  ptxas contracts mul+add pairs -- including packed mul.rn.f32x2 /
    add.rn.f32x2 even with explicit .rn -- so the build is checked: no FFMA / FFMA2 / HFMA-free
    distance kernels only (a contracted FMA would silently break bit-exactness)."""
    import shutil
    import subprocess
    import pytest
    from pocketsphinx_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    fn, bad = None, []
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif fn and any(k in fn for k in ("topn", "ms_dist")) and "FFMA" in line:
            bad.append((fn[:60], line.strip()[:80]))
    assert not bad, bad[:5]
    assert "FMUL2" in sass and "FADD2" in sass, "packed FP32 path missing from the build"
