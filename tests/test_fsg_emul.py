"""CPU: the phase code of the device grammar search (pocketsphinx_b200/csrc/psb_fsg_core.h -- the very
source fsg_search_kernel is compiled from) built for the host by tests/emul/fsg_emul.cpp and run one
"thread" at a time, in ascending and in descending thread order, against the reference's golden
history tables (tests/golden/en_us_fsg.npz).  Both orders must reproduce every row: the closed forms
that replace the reference's list walks are right, and no phase depends on the order its threads
run in.  (What this cannot show -- barriers, the block scan, memory spaces -- is what the -m gpu
test of psb_fsg_batch_device is for.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden

TAGS = ("go", "go_hmmpf", "cmd", "cmd_wide", "cmd_hmmpf")
ARGT = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
        C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module", params=["ascending", "descending"])
def emul(request, tmp_path_factory):
    from oracle import oracle
    oracle.build()
    out = str(tmp_path_factory.mktemp("fsgemul") / ("libfsgemul_%s.so" % request.param))
    odir = os.path.dirname(oracle.LIB_PATH)
    cmd = ["g++", "-O1", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror"]
    if request.param == "descending":
        cmd.append("-DPSB_FSG_EMUL_REVERSE")
    cmd += ["-o", out, os.path.join(ROOT, "tests", "emul", "fsg_emul.cpp"), "-L" + odir, "-lpsoracle", "-Wl,-rpath," + odir]
    subprocess.check_call(cmd)
    f = C.CDLL(out).fsg_emul_run
    f.restype = C.c_int32
    f.argtypes = ARGT
    return f


def _run(f, m, c, scr, cap):
    tp = np.ascontiguousarray(m["tp"], np.uint8)
    sseq = np.ascontiguousarray(m["sseq"], np.uint16)
    a = {k: np.ascontiguousarray(c[k], np.int32) for k in ("pnodes", "roots", "links", "nulloff", "nullarc")}
    scr = np.ascontiguousarray(scr, np.int16)
    hist = np.zeros((cap, 13), np.int32)
    n = f(tp.shape[1], _p(tp), _p(sseq), len(a["pnodes"]), _p(a["pnodes"]), len(a["roots"]), _p(a["roots"]),
          len(a["links"]), _p(a["links"]), _p(a["nulloff"]), _p(a["nullarc"]), int(c["n_ciphone"]), int(c["silcipid"]),
          int(c["start_state"]), int(c["beam"]), int(c["pbeam"]), int(c["wbeam"]), int(c["maxhmmpf"]), _p(scr),
          scr.shape[1], scr.shape[0], _p(hist), cap)
    return hist[:max(0, min(n, cap))], n


def _case(g, tag):
    return {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + ".")}


@pytest.mark.parametrize("tag", TAGS)
def test_phase_code_reproduces_reference_history(emul, tag):
    m, scr = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")["senscr"]
    c = _case(golden("en_us_fsg.npz"), tag)
    hist, n = _run(emul, m, c, scr, len(c["hist"]) + 16)
    assert n == len(c["hist"]) and np.array_equal(hist, c["hist"])


def test_phase_code_truncated_table_and_short_utterances(emul):
    from oracle import oracle
    m, scr = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")["senscr"]
    c = _case(golden("en_us_fsg.npz"), "cmd")
    hist, n = _run(emul, m, c, scr, 100)                       # rows past cap are dropped, the count is not
    assert n == len(c["hist"]) and np.array_equal(hist, c["hist"][:100])
    for T in (0, 1, 7, 120):
        want = oracle.fsg_run(m["tp"], m["sseq"], c, scr[:T])
        hist, n = _run(emul, m, c, scr[:T], len(want) + 4)
        assert n == len(want) and np.array_equal(hist, want), T


def test_graph_validation_rejects_non_trees(emul):
    m, scr = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")["senscr"]
    c = dict(_case(golden("en_us_fsg.npz"), "go"))
    pn = c["pnodes"].copy()
    inner = np.nonzero((pn[:, 7] == 0) & (pn[:, 2] >= 0))[0]
    a, b = inner[0], inner[1]
    pn[b, 2] = pn[a, 2]                                         # two parents share a child chain
    c["pnodes"] = pn
    assert _run(emul, m, c, scr[:5], 64)[1] == -1
    c["pnodes"] = _case(golden("en_us_fsg.npz"), "go")["pnodes"].copy()
    c["pnodes"][0, 3] = 10 ** 6                                 # sibling out of range
    assert _run(emul, m, c, scr[:5], 64)[1] == -1
