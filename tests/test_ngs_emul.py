"""CPU: the phase code of the device first pass (pocketsphinx_b200/csrc/psb_ngs_core.h -- the source
ngs_fwdtree_kernel is compiled from) built for the host by tests/emul/ngs_emul.cpp and run one
"thread" at a time, ascending and descending, against the reference's golden backpointer tables
(tests/golden/en_us_fwdtree.npz): every bp_table row, the right-context score stack and
bp_table_idx, for default / wide / narrow beams, -maxwpf, -maxhmmpf, penalties and look-ahead."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden

TAGS = ("default", "wide", "narrow", "maxwpf", "abs", "pen", "lookahead")
ARGT = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
        C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@pytest.fixture(scope="module", params=["ascending", "descending"])
def emul(request, tmp_path_factory):
    from oracle import oracle
    oracle.build()
    out = str(tmp_path_factory.mktemp("ngsemul") / ("libngsemul_%s.so" % request.param))
    odir = os.path.dirname(oracle.LIB_PATH)
    cmd = ["g++", "-O1", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror"]
    if request.param == "descending":
        cmd.append("-DPSB_FSG_EMUL_REVERSE")
    cmd += ["-o", out, os.path.join(ROOT, "tests", "emul", "ngs_emul.cpp"), "-L" + odir, "-lpsoracle", "-Wl,-rpath," + odir]
    subprocess.check_call(cmd)
    f = C.CDLL(out).ngs_emul_run
    f.restype = C.c_int32
    f.argtypes = ARGT
    return f


def run_emul(f, m, info, model, scr, bp_cap, bss_cap, pl_pen=None, pl_window=0, lm_arrays=None):
    tp = np.ascontiguousarray(m["tp"], np.uint8)
    sseq = np.ascontiguousarray(m["sseq"], np.uint16)
    info = np.ascontiguousarray(info, np.int32)
    model = np.ascontiguousarray(model, np.int32)
    cit = np.ascontiguousarray(m["phone_tmat"][:int(info[6])], np.int32)
    scr = np.ascontiguousarray(scr, np.int16)
    T = len(scr)
    pen = None
    if pl_pen is not None and pl_window > 0 and T > 0:
        pen = np.ascontiguousarray(np.asarray(pl_pen, np.int32)[:T])       # the phone loop's own table; the window is applied inside
    lma = None if lm_arrays is None else np.ascontiguousarray(lm_arrays, np.int32)
    bp = np.zeros((bp_cap, 10), np.int32)
    bss = np.zeros(bss_cap, np.int32)
    idx = np.zeros(T + 2, np.int32)
    bn = C.c_int32()
    n = f(tp.shape[1], _p(tp), tp.shape[0], _p(sseq), len(sseq), _p(cit), _p(info), _p(model), len(model), _p(lma), 0 if lma is None else len(lma), _p(scr), scr.shape[1], T,
          _p(pen), int(pl_window), _p(bp), bp_cap, _p(bss), bss_cap, C.byref(bn), _p(idx))
    return n, bp[:max(n, 0)], bss[:bn.value if n >= 0 else 0], idx[:T + 1]


def _case(g, tag):
    return {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + ".")}


@pytest.mark.parametrize("tag", TAGS)
def test_phase_code_reproduces_reference_bptable(emul, tag):
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    c = _case(golden("en_us_fwdtree.npz"), tag)
    la = dict(pl_pen=gf["pl_pen"], pl_window=int(gf["pl_params"][4])) if tag == "lookahead" else {}
    n, bp, bss, idx = run_emul(emul, m, c["info"], c["model"], gf["senscr"], len(c["bp"]) + 8, len(c["bss"]) + 64, **la)
    assert n == len(c["bp"]) and np.array_equal(bp, c["bp"])
    assert np.array_equal(bss, c["bss"]) and np.array_equal(idx, c["bp_idx"])


def test_phase_code_short_utterances_and_full_tables(emul):
    from oracle import oracle
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    c = _case(golden("en_us_fwdtree.npz"), "default")
    scr = gf["senscr"]
    for T in (0, 1, 3, 40, 150):
        want = oracle.fwdtree_run(m["tp"], m["sseq"], m["phone_tmat"][:int(c["info"][6])], c["info"], c["model"], scr[:T])
        n, bp, bss, idx = run_emul(emul, m, c["info"], c["model"], scr[:T], len(want[0]) + 4, len(want[1]) + 64)
        assert n == len(want[0]) and np.array_equal(bp, want[0]) and np.array_equal(bss, want[1]), T
        assert np.array_equal(idx, want[2]), T
    assert run_emul(emul, m, c["info"], c["model"], scr, 100, 100000)[0] == -2            # table full: an error, not a truncation


def test_graph_validation(emul):
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    c = _case(golden("en_us_fwdtree.npz"), "default")
    model = c["model"].copy()
    n_root = int(c["info"][2])
    model[n_root * 5 + 5] = 0                                   # non-root channel 0: alt -> itself/another: two parents
    model[n_root * 5 + 4] = 0
    assert run_emul(emul, m, c["info"], model, gf["senscr"][:5], 64, 4096)[0] == -1
    assert run_emul(emul, m, c["info"], c["model"][:1000], gf["senscr"][:5], 64, 4096)[0] == -1      # block shorter than info says


@pytest.mark.parametrize("blocks,ok", [("60", True), ("30", True), ("6", False)])
def test_fanout_block_pool_reuse_and_exhaustion(emul, monkeypatch, blocks, ok):
    """The right-context fan-out lives in a per-utterance pool of blocks (a word owns one while any of its
    channels is allocated).  Fewer blocks than multi-phone words (about 100 here) forces reuse: same tables;
    a pool that runs dry is an error, not a wrong answer."""
    monkeypatch.setenv("PSB_NGS_BLOCKS", blocks)
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    c = _case(golden("en_us_fwdtree.npz"), "default")
    n, bp, bss, idx = run_emul(emul, m, c["info"], c["model"], gf["senscr"], len(c["bp"]) + 8, len(c["bss"]) + 64)
    if ok:
        assert n == len(c["bp"]) and np.array_equal(bp, c["bp"]) and np.array_equal(bss, c["bss"]) and np.array_equal(idx, c["bp_idx"])
    else:
        assert n == -4
