"""CPU: the phase code of the device second pass (pocketsphinx_b200/csrc/psb_ngf_core.h -- the source
ngs_fwdflat_kernel is compiled from) built for the host by tests/emul/ngf_emul.cpp, run one
"thread" at a time in both orders behind the host-emulated FIRST pass (tests/emul/ngs_emul.cpp), against
the reference's golden second-pass backpointer tables (tests/golden/en_us_fwdtree.npz, flat_*)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from test_ngs_emul import run_emul as run_first

TAGS = ("flat_default", "flat_wide", "flat_narrow")
ARGT = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
        C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _build(tmp, name, src, reverse):
    from oracle import oracle
    oracle.build()
    out = str(tmp / ("lib%s_%s.so" % (name, "rev" if reverse else "fwd")))
    odir = os.path.dirname(oracle.LIB_PATH)
    cmd = ["g++", "-O1", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror", "-ffp-contract=off"]
    if reverse:
        cmd.append("-DPSB_FSG_EMUL_REVERSE")
    cmd += ["-o", out, os.path.join(ROOT, "tests", "emul", src), "-L" + odir, "-lpsoracle", "-Wl,-rpath," + odir]
    subprocess.check_call(cmd)
    return C.CDLL(out)


@pytest.fixture(scope="module", params=["ascending", "descending"])
def emuls(request, tmp_path_factory):
    from test_ngs_emul import ARGT as ARGT1
    tmp = tmp_path_factory.mktemp("ngfemul")
    rev = request.param == "descending"
    f1 = _build(tmp, "ngsemul", "ngs_emul.cpp", rev).ngs_emul_run
    f1.restype = C.c_int32
    f1.argtypes = ARGT1
    f2 = _build(tmp, "ngfemul", "ngf_emul.cpp", rev).ngf_emul_run
    f2.restype = C.c_int32
    f2.argtypes = ARGT
    return f1, f2


def run_second(f, m, info, model, bp1, scr, bp_cap, bss_cap, lm_arrays=None):
    tp = np.ascontiguousarray(m["tp"], np.uint8)
    sseq = np.ascontiguousarray(m["sseq"], np.uint16)
    info = np.ascontiguousarray(info, np.int32)
    model = np.ascontiguousarray(model, np.int32)
    nci = int(info[6])
    cit = np.ascontiguousarray(m["phone_tmat"][:nci], np.int32)
    cis = np.ascontiguousarray(m["phone_ssid"][:nci], np.int32)
    n1 = -1 if bp1 is None else len(bp1)                        # None: no first pass (-fwdtree no)
    bp1 = np.zeros((1, 10), np.int32) if bp1 is None else np.ascontiguousarray(bp1, np.int32)
    scr = np.ascontiguousarray(scr, np.int16)
    T = len(scr)
    lma = None if lm_arrays is None else np.ascontiguousarray(lm_arrays, np.int32)
    bp = np.zeros((bp_cap, 10), np.int32)
    bss = np.zeros(bss_cap, np.int32)
    idx = np.zeros(T + 2, np.int32)
    bn = C.c_int32()
    n = f(tp.shape[1], _p(tp), tp.shape[0], _p(sseq), len(sseq), _p(cit), _p(cis), _p(info), _p(model), len(model), _p(lma), 0 if lma is None else len(lma), _p(bp1), n1,
          _p(scr), scr.shape[1], T, _p(bp), bp_cap, _p(bss), bss_cap, C.byref(bn), _p(idx))
    return n, bp[:max(n, 0)], bss[:bn.value if n >= 0 else 0], idx[:T + 1]


def _case(g, tag):
    return {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + ".")}


@pytest.mark.parametrize("tag", TAGS)
def test_both_passes_of_phase_code_reproduce_reference(emuls, tag):
    f1, f2 = emuls
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    c = _case(golden("en_us_fwdtree.npz"), tag)
    scr = gf["senscr"]
    la = dict(pl_pen=gf["pl_pen"], pl_window=int(gf["pl_params"][4])) if tag == "flat_default" else {}
    n1, bp1, _, _ = run_first(f1, m, c["info"], c["model"], scr, 8192, 1 << 18, **la)
    assert n1 > 0
    n, bp, bss, idx = run_second(f2, m, c["info"], c["model"], bp1, scr, len(c["bp"]) + 8, len(c["bss"]) + 64)
    assert n == len(c["bp"]) and np.array_equal(bp, c["bp"])
    assert np.array_equal(bss, c["bss"]) and np.array_equal(idx, c["bp_idx"])


def test_second_pass_short_utterances_and_full_tables(emuls):
    from oracle import oracle
    f1, f2 = emuls
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    c = _case(golden("en_us_fwdtree.npz"), "flat_wide")
    nci = int(c["info"][6])
    for T in (0, 1, 9, 60, 170):
        scr = gf["senscr"][:T]
        bp1 = oracle.fwdtree_run(m["tp"], m["sseq"], m["phone_tmat"][:nci], c["info"], c["model"], scr)[0]
        want = oracle.fwdflat_run(m["tp"], m["sseq"], m["phone_tmat"][:nci], m["phone_ssid"][:nci], c["info"], c["model"], bp1, scr)
        n, bp, bss, idx = run_second(f2, m, c["info"], c["model"], bp1, scr, len(want[0]) + 4, len(want[1]) + 64)
        assert n == len(want[0]) and np.array_equal(bp, want[0]) and np.array_equal(bss, want[1]), T
        assert np.array_equal(idx, want[2]), T
    scr = gf["senscr"]
    bp1 = oracle.fwdtree_run(m["tp"], m["sseq"], m["phone_tmat"][:nci], c["info"], c["model"], scr)[0]
    assert run_second(f2, m, c["info"], c["model"], bp1, scr, 50, 100000)[0] == -2


@pytest.mark.parametrize("model_key", ["flat_default", "nodense"])
def test_both_passes_with_the_array_lm(emuls, model_key):
    """Trigram scores from the LM as arrays (psb_lm_core.h) instead of the dense table -- with the table
    still in the model block, and with a block exported without it (info[26] = 0)."""
    f1, f2 = emuls
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    g = golden("en_us_fwdtree.npz")
    want = _case(g, "flat_default")
    info, model = g[model_key + ".info"], g[model_key + ".model"]
    if model_key == "nodense":
        assert int(info[26]) == 0 and len(model) < len(want["model"]) // 3
    la = dict(pl_pen=gf["pl_pen"], pl_window=int(gf["pl_params"][4]))
    n1, bp1, _, _ = run_first(f1, m, info, model, gf["senscr"], 8192, 1 << 18, lm_arrays=g["lmarr"], **la)
    assert n1 > 0 and np.array_equal(bp1, _case(g, "lookahead")["bp"])
    n, bp, bss, idx = run_second(f2, m, info, model, bp1, gf["senscr"], len(want["bp"]) + 8, len(want["bss"]) + 64, lm_arrays=g["lmarr"])
    assert n == len(want["bp"]) and np.array_equal(bp, want["bp"]) and np.array_equal(bss, want["bss"]) and np.array_equal(idx, want["bp_idx"])
    lma = g["lmarr"].copy()
    lma[10 + int(lma[7]) + 2 * int(lma[1]) + 1] = 10 ** 6       # uni_next[1] beyond the bigram array
    assert run_first(f1, m, info, model, gf["senscr"][:5], 64, 4096, lm_arrays=lma)[0] == -1


@pytest.mark.parametrize("channels,ok", [("1500", True), ("300", False)])
def test_state_area_holds_the_utterance_vocabulary_only(emuls, monkeypatch, channels, ok):
    """Second-pass channels are laid out per utterance over its vocabulary (about 60 of the 110 words here),
    not over every LM word: a state area smaller than the whole LM's chains still gives the same tables; one
    that cannot hold the vocabulary is an error."""
    from oracle import oracle
    f1, f2 = emuls
    monkeypatch.setenv("PSB_NGF_CHANNELS", channels)
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    c = _case(golden("en_us_fwdtree.npz"), "flat_wide")
    bp1 = oracle.fwdtree_run(m["tp"], m["sseq"], m["phone_tmat"][:int(c["info"][6])], c["info"], c["model"], gf["senscr"])[0]
    n, bp, bss, idx = run_second(f2, m, c["info"], c["model"], bp1, gf["senscr"], len(c["bp"]) + 8, len(c["bss"]) + 64)
    if ok:
        assert n == len(c["bp"]) and np.array_equal(bp, c["bp"]) and np.array_equal(bss, c["bss"]) and np.array_equal(idx, c["bp_idx"])
    else:
        assert n == -4
