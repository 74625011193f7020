"""CPU (needs oracle/_ref/libpsref.so): pocketsphinx_b200.lextree.build_ngram_search -- the flattened n-gram
search (info + every model section: root / interior channels, word table with homophone chains, single-phone
words, dict2pid tables, LM membership, pronunciations with their word-internal senone sequences) built from the
FILES ALONE must equal what the maintainer-side binding exports from a reference decoder: demo LM, tidigits,
and cmudict + the 72 k-word en-us LM (723 roots, 152 500 interior channels); other beams and penalties."""
import os

import numpy as np
import pytest

from oracle import refdrv
from pocketsphinx_b200 import lextree

pytestmark = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")
REF = os.path.dirname(refdrv.LIB_PATH)
EN, TD = os.path.join(REF, "model", "en-us"), os.path.join(REF, "model", "tidigits_hmm")
CASES = {"turtle": (EN, os.path.join(REF, "data", "turtle.lm.bin"), os.path.join(REF, "data", "turtle.dic")),
         "tidigits": (TD, os.path.join(REF, "model", "tidigits_lm", "tidigits.lm.bin"), os.path.join(REF, "model", "tidigits_lm", "tidigits.dic")),
         "cmudict": (EN, os.path.join(REF, "model", "en-us.lm.bin"), os.path.join(REF, "model", "cmudict-en-us.dict"))}
RESULT_SLOTS = {0, 24, 25, 27}                                   # frames, table sizes and score of the driver's decode


def from_files(hd, lm, dic, **kv):
    g = lextree.ngram_search_from_files(hd, dic, lm, **kv)
    return g["info"], g["model"], g["lm_arrays"]


@pytest.mark.parametrize("name,kv", [("turtle", {}), ("tidigits", {}), ("cmudict", {}),
                                     ("turtle", dict(beam="1e-60", wbeam="1e-40", pbeam="1e-55", maxwpf="5", maxhmmpf="500", lw="9.5",
                                                     fwdflatlw="7", silprob="0.02", fillprob="1e-5", pip="0.5", nwpen="0.8",
                                                     fwdflatbeam="1e-50", fwdflatefwid="3", fwdflatsfwin="20"))],
                         ids=["turtle", "tidigits", "cmudict-72k", "turtle-settings"])
def test_search_description_equals_the_exported_one(name, kv):
    hd, lm, dic = CASES[name]
    if not os.path.exists(lm):
        pytest.skip("LM file not present")
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    r = refdrv.fwdtree(hd, lm, dic, pcm, dense_lm=False, fwdflat="yes", **kv)
    info, model, arr = from_files(hd, lm, dic, **kv)
    assert [i for i in range(40) if i not in RESULT_SLOTS and info[i] != r["info"][i]] == []
    assert model.shape == r["model"].shape and np.array_equal(model, r["model"])
    assert np.array_equal(arr, refdrv.lm_arrays(hd, lm, dic, **kv)[0])
    if name == "cmudict":
        assert info[2] > 700 and info[3] > 150_000


def test_files_alone_decode_like_the_reference():
    """End to end on the CPU side: description + LM block from the files, the oracle's first and second pass on the
    reference's senone scores -> the reference's own backpointer tables and hypothesis."""
    from oracle import oracle
    from pocketsphinx_b200 import api
    hd, lm, dic = CASES["turtle"]
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    want = refdrv.fwdtree(hd, lm, dic, pcm, dense_lm=False, fwdflat="yes")
    first = refdrv.fwdtree(hd, lm, dic, pcm, dense_lm=False)
    info, model, arr = from_files(hd, lm, dic)
    ref = refdrv.RefModel(hd)
    pk = ref.packed()
    scr = np.ascontiguousarray(ref.score(ref.featurize_fresh(pcm)))
    ref.close()
    nc = int(info[6])
    o1 = oracle.fwdtree_run(pk["tp"], pk["sseq"], pk["phone_tmat"][:nc], info, model, scr, lm_arrays=arr)
    assert np.array_equal(o1[0], first["bp"]) and np.array_equal(o1[1], first["bss"])
    o2 = oracle.fwdflat_run(pk["tp"], pk["sseq"], pk["phone_tmat"][:nc], pk["phone_ssid"][:nc], info, model, o1[0], scr, lm_arrays=arr)
    assert np.array_equal(o2[0], want["bp"]) and np.array_equal(o2[1], want["bss"]) and np.array_equal(o2[2], want["bp_idx"])
    g = lextree.ngram_search_from_files(hd, dic, lm)
    words = g["words"]
    assert np.array_equal(g["ci_tmat"], pk["phone_tmat"][:nc]) and np.array_equal(g["ci_ssid"], pk["phone_ssid"][:nc])
    entry, score, seg = api.ngram_hyp(o2[0], o2[2], len(scr), int(info[20]))
    real = [words[w].split("(")[0] for w in seg[:, 1] if not (int(info[22]) <= w <= int(info[23]))]
    assert " ".join(real) == want["hyp"] == "go forward ten meters"


def test_files_alone_equals_the_golden_description_the_gpu_tests_use():
    """tests/golden/en_us_fwdtree.npz (nodense.info / nodense.model / lmarr) is what the gated GPU tests hand to the
    kernels with the array LM; the same blocks come out of the files, so those tests also cover this path."""
    from conftest import golden
    g = golden("en_us_fwdtree.npz")
    hd, lm, dic = CASES["turtle"]
    d = lextree.ngram_search_from_files(hd, dic, lm)
    assert [i for i in range(40) if i not in RESULT_SLOTS and d["info"][i] != g["nodense.info"][i]] == []
    assert np.array_equal(d["model"], g["nodense.model"]) and np.array_equal(d["lm_arrays"], g["lmarr"])


@pytest.mark.parametrize("seed", [0, 1])
def test_random_sub_dictionaries_against_the_export(tmp_path, seed):
    """4000 random cmudict lines (odd seeds: shuffled, so alternates can precede or lose their base word and are
    dropped like dict_add_word drops them), random beam and weights, the 72 k-word LM: description and LM block from
    the files equal the export."""
    import random
    hd, lm, big = CASES["cmudict"]
    if not os.path.exists(lm):
        pytest.skip("LM file not present")
    rng = random.Random(seed)
    lines = open(big, encoding="latin-1").read().split("\n")
    sub = [lines[i] for i in sorted(rng.sample(range(len(lines)), 4000)) if lines[i].strip()]
    if seed % 2:
        rng.shuffle(sub)
    p = str(tmp_path / "sub.dic")
    open(p, "w", encoding="latin-1").write("\n".join(sub) + "\n")
    kv = dict(beam="1e-%d" % rng.randint(30, 80), lw=str(rng.choice([5, 6.5, 9.5])), wip=str(rng.choice([0.2, 0.65])))
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)[:8000]
    r = refdrv.fwdtree(hd, lm, p, pcm, dense_lm=False, fwdflat="yes", **kv)
    g = lextree.ngram_search_from_files(hd, p, lm, **kv)
    assert [i for i in range(40) if i not in RESULT_SLOTS and g["info"][i] != r["info"][i]] == []
    assert np.array_equal(g["model"], r["model"]) and np.array_equal(g["lm_arrays"], refdrv.lm_arrays(hd, lm, p, **kv)[0])
    assert len(g["words"]) < len(sub) + 10 and int(g["info"][3]) > 5000
