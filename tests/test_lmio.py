"""CPU: pocketsphinx_b200.lmio -- a binary trie LM read WITHOUT the reference must give, word for word, the
int32 block the maintainer-side binding unpacks from the reference's loaded trie (cuda_ngram_export_lm):
the reference's trigram demo LM, its bigram tidigits LM, other language weights, and its 72 k-word en-us
LM against the 134 865-word cmudict (bit-packed arrays, 16-bit quantised probabilities, <UNK> / unknown
words, alternate pronunciations)."""
import os

import numpy as np
import pytest

from conftest import golden
from oracle import refdrv
from pocketsphinx_b200 import lmio

REF = os.path.dirname(refdrv.LIB_PATH)
live = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")
EN, TD = os.path.join(REF, "model", "en-us"), os.path.join(REF, "model", "tidigits_hmm")
TURTLE = (EN, os.path.join(REF, "data", "turtle.lm.bin"), os.path.join(REF, "data", "turtle.dic"))
DIGITS = (TD, os.path.join(REF, "model", "tidigits_lm", "tidigits.lm.bin"), os.path.join(REF, "model", "tidigits_lm", "tidigits.dic"))
BIG = (EN, os.path.join(REF, "model", "en-us.lm.bin"), os.path.join(REF, "model", "cmudict-en-us.dict"))


@live
@pytest.mark.parametrize("case,kv", [(TURTLE, {}), (TURTLE, dict(lw="9.5", wip="0.2")), (DIGITS, {}), (BIG, {})],
                         ids=["turtle", "turtle-weights", "tidigits-bigram", "en-us-72k"])
def test_lm_arrays_equal_the_bindings(case, kv):
    hd, lm, dic = case
    if not os.path.exists(lm):
        pytest.skip("LM file not present")
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    want, _ = refdrv.lm_arrays(hd, lm, dic, **kv)
    vocab = refdrv.fwdtree(hd, lm, dic, pcm, dense_lm=False, **kv)["vocab"]       # the dictionary's word strings by id
    L = lmio.read_lm_bin(lm)
    got = lmio.lm_arrays(L, vocab, lw=float(kv.get("lw", 6.5)), wip=float(kv.get("wip", 0.65)))
    assert got.dtype == np.int32 and got.shape == want.shape and np.array_equal(got, want)
    assert got[7] == len(vocab) and got[1] == len(L["words"])
    if case is BIG:
        unk = L["words"].index("<UNK>") if "<UNK>" in L["words"] else -1
        assert L["counts"][1] > 2_000_000 and (got[10:10 + len(vocab)] == unk).sum() > 50_000    # dictionary words the LM lacks


def test_golden_arrays_from_the_file_alone():
    """The LM block the gated GPU tests hand to the kernels (tests/golden/en_us_fwdtree.npz: lmarr) from the file alone."""
    g = golden("en_us_fwdtree.npz")
    if not os.path.exists(TURTLE[1]):
        pytest.skip("turtle.lm.bin not present")
    vocab = str(g["default.vocab"]).split("\n")
    got = lmio.lm_arrays(lmio.read_lm_bin(TURTLE[1]), vocab)
    assert np.array_equal(got, g["lmarr"])


def test_damaged_files_are_errors(tmp_path):
    if not os.path.exists(TURTLE[1]):
        pytest.skip("turtle.lm.bin not present")
    raw = open(TURTLE[1], "rb").read()
    for name, data in (("magic", b"Tree" + raw[4:]), ("short", raw[:len(raw) // 2]), ("order", raw[:19] + b"\x07" + raw[20:]),
                       ("strings", raw[:-40])):
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        with pytest.raises((ValueError, NotImplementedError, IndexError, __import__("struct").error)):
            lmio.lm_arrays(lmio.read_lm_bin(p), ["go"])


@live
@pytest.mark.parametrize("case", [TURTLE, DIGITS, BIG], ids=["turtle", "tidigits", "cmudict-134k"])
def test_dictionary_in_the_references_word_id_order(case):
    """read_dict: word strings, base word ids, filler range and first / last phones of every dictionary id as
    the reference's dict_t has them (main file, the model's noisedict, then <s> </s> <sil>), so that the LM
    block can be built from the files alone."""
    from pocketsphinx_b200 import s3io
    hd, lm, dic = case
    if not os.path.exists(lm):
        pytest.skip("LM file not present")
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    r = refdrv.fwdtree(hd, lm, dic, pcm, dense_lm=False)
    md = s3io.read_mdef(os.path.join(hd, "mdef"))
    nd = os.path.join(hd, "noisedict")
    words, prons, base, filler_start = lmio.read_dict(dic, nd if os.path.exists(nd) else None, md["ciname"])
    assert words == r["vocab"] and np.array_equal(base, r["words"][:, 5])
    assert (filler_start, len(words) - 1) == (int(r["info"][22]), int(r["info"][23]))
    ci = {n: i for i, n in enumerate(md["ciname"])}
    assert np.array_equal([ci[p[0]] for p in prons], r["words"][:, 0]) and np.array_equal([ci[p[-1]] for p in prons], r["words"][:, 1])
    assert np.array_equal([len(p) == 1 for p in prons], r["words"][:, 3] != 0)
    want, _ = refdrv.lm_arrays(hd, lm, dic)
    assert np.array_equal(lmio.lm_arrays(lmio.read_lm_bin(lm), words), want)          # files alone -> the kernels' LM block


def test_dictionary_rules(tmp_path):
    d = tmp_path / "d.dic"
    d.write_text(";; comment\n## another\nhello HH AH L OW\nhello(2) HH EH L OW\nworld W ER L D\nhello HH AH\n"
                 "orphan(2) AO R\nnopron\nzz ZZ\n\nbye B AY\n")
    f = tmp_path / "noise"
    f.write_text("<sil> SIL\n+noise+ +NSN+\n")
    ci = ["HH", "AH", "L", "OW", "EH", "W", "ER", "D", "B", "AY", "SIL", "+NSN+", "AO", "R"]
    words, prons, base, fs = lmio.read_dict(str(d), str(f), ci)
    assert words == ["hello", "hello(2)", "world", "bye", "<sil>", "+noise+", "<s>", "</s>"]      # duplicate, orphan alternate,
    assert base.tolist() == [0, 0, 2, 3, 4, 5, 6, 7] and fs == 4                                    # unknown phone, no pronunciation dropped
    d.write_text("<s> SIL\n")
    with pytest.raises(ValueError):
        lmio.read_dict(str(d), None, ci)
