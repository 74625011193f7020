"""CPU: data-race check of the device searches' phase code.  The three host harnesses are rebuilt with
-DPSB_FSG_RACECHECK (tests/emul/psb_fsg_racecheck.h): every access to an utterance's
mutable state and to the block-shared scalars is recorded with (phase, thread), phases ending at
FSG_SYNC(); a location touched by two different threads in one phase with at least one (value-changing)
write is reported.  All fixtures must run clean AND still reproduce the reference's tables; one build
with a barrier deliberately removed (the one a read-through had to add by hand) must be reported, so
that a silent detector cannot pass."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import test_fsg_emul as TF
import test_ngf_emul as TG
import test_ngs_emul as TN
from conftest import ROOT, golden


def _build(tmp, src, extra=()):
    from oracle import oracle
    oracle.build()
    out = str(tmp / (src.replace(".cpp", "") + "_" + "_".join(x.strip("-D") for x in extra) + "race.so"))
    odir = os.path.dirname(oracle.LIB_PATH)
    subprocess.check_call(["g++", "-O1", "-fPIC", "-shared", "-Wall", "-Wextra", "-ffp-contract=off", "-DPSB_FSG_RACECHECK", "-I" + os.path.join(ROOT, "tests", "emul"), *extra,
                           "-o", out, os.path.join(ROOT, "tests", "emul", src), "-L" + odir, "-lpsoracle", "-Wl,-rpath," + odir])
    L = C.CDLL(out)
    L.emul_race_count.restype = C.c_long
    return L


@pytest.fixture(scope="module")
def tmp(tmp_path_factory):
    return tmp_path_factory.mktemp("racecheck")


def test_grammar_search_phase_code_is_race_free(tmp):
    L = _build(tmp, "fsg_emul.cpp")
    f = L.fsg_emul_run
    f.restype = C.c_int32
    f.argtypes = TF.ARGT
    m, scr = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")["senscr"]
    for tag in TF.TAGS:
        c = TF._case(golden("en_us_fsg.npz"), tag)
        hist, n = TF._run(f, m, c, scr, len(c["hist"]) + 16)
        assert n == len(c["hist"]) and np.array_equal(hist, c["hist"]), tag
    assert L.emul_race_count() == 0


def test_first_pass_phase_code_is_race_free(tmp):
    L = _build(tmp, "ngs_emul.cpp")
    f = L.ngs_emul_run
    f.restype = C.c_int32
    f.argtypes = TN.ARGT
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    for tag in TN.TAGS:
        c = TN._case(golden("en_us_fwdtree.npz"), tag)
        la = dict(pl_pen=gf["pl_pen"], pl_window=int(gf["pl_params"][4])) if tag == "lookahead" else {}
        n, bp, bss, idx = TN.run_emul(f, m, c["info"], c["model"], gf["senscr"], len(c["bp"]) + 8, len(c["bss"]) + 64, **la)
        assert n == len(c["bp"]) and np.array_equal(bp, c["bp"]) and np.array_equal(bss, c["bss"]), tag
    assert L.emul_race_count() == 0


def test_second_pass_phase_code_is_race_free(tmp):
    from oracle import oracle
    L = _build(tmp, "ngf_emul.cpp")
    f = L.ngf_emul_run
    f.restype = C.c_int32
    f.argtypes = TG.ARGT
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    scr = gf["senscr"]
    for tag in TG.TAGS:
        c = TG._case(golden("en_us_fwdtree.npz"), tag)
        la = dict(pl_pen=gf["pl_pen"], pl_window=int(gf["pl_params"][4])) if tag == "flat_default" else {}
        bp1 = oracle.fwdtree_run(m["tp"], m["sseq"], m["phone_tmat"][:int(c["info"][6])], c["info"], c["model"], scr, **la)[0]
        n, bp, bss, idx = TG.run_second(f, m, c["info"], c["model"], bp1, scr, len(c["bp"]) + 8, len(c["bss"]) + 64)
        assert n == len(c["bp"]) and np.array_equal(bp, c["bp"]) and np.array_equal(bss, c["bss"]), tag
    assert L.emul_race_count() == 0


def test_detector_reports_a_removed_barrier(tmp):
    L = _build(tmp, "ngs_emul.cpp", ("-DNGS_TEST_INJECT_RACE",))
    f = L.ngs_emul_run
    f.restype = C.c_int32
    f.argtypes = TN.ARGT
    m, gf = golden("en_us_ptm_model.npz"), golden("en_us_goforward.npz")
    c = TN._case(golden("en_us_fwdtree.npz"), "default")
    TN.run_emul(f, m, c["info"], c["model"], gf["senscr"][:40], 4096, 1 << 16)
    assert L.emul_race_count() >= 30                            # once per frame: all threads read bpidx, the leader moves it
