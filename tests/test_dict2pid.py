"""CPU (needs oracle/_ref/libpsref.so): pocketsphinx_b200.dict2pid -- the cross-word triphone tables built from
the files alone (binary mdef's triphone tree + dictionary) must equal the `rs_n | rs_ssid | rs_cimap | ldiph_lc`
sections the maintainer-side binding exports from the reference's own dict2pid_t: demo dictionary, tidigits
(its own phone set, 5-state model) and the 134 865-word cmudict."""
import os

import numpy as np
import pytest

from oracle import refdrv
from pocketsphinx_b200 import dict2pid, lmio, s3io

pytestmark = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")
REF = os.path.dirname(refdrv.LIB_PATH)
EN, TD = os.path.join(REF, "model", "en-us"), os.path.join(REF, "model", "tidigits_hmm")
CASES = {"turtle": (EN, os.path.join(REF, "data", "turtle.lm.bin"), os.path.join(REF, "data", "turtle.dic")),
         "tidigits": (TD, os.path.join(REF, "model", "tidigits_lm", "tidigits.lm.bin"), os.path.join(REF, "model", "tidigits_lm", "tidigits.dic")),
         "cmudict": (EN, os.path.join(REF, "model", "en-us.lm.bin"), os.path.join(REF, "model", "cmudict-en-us.dict"))}


@pytest.mark.parametrize("name", list(CASES))
def test_tables_equal_the_references(name):
    hd, lm, dic = CASES[name]
    if not os.path.exists(lm):
        pytest.skip("LM file not present")
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    r = refdrv.fwdtree(hd, lm, dic, pcm, dense_lm=False)
    md = s3io.read_mdef(os.path.join(hd, "mdef"))
    nd = os.path.join(hd, "noisedict")
    words, prons, base, fs = lmio.read_dict(dic, nd if os.path.exists(nd) else None, md["ciname"])
    ci = {n: i for i, n in enumerate(md["ciname"])}
    got = dict2pid.build(md, [[ci[x] for x in p] for p in prons])
    info, model = r["info"], r["model"]
    nc = int(info[6])
    o = int(info[2]) * 5 + int(info[3]) * 6 + int(info[1]) * 8 + int(info[4]) * 5
    for key, size, shape in (("rs_n", nc * nc, (nc, nc)), ("rs_ssid", nc ** 3, (nc, nc, nc)), ("rs_cimap", nc ** 3, (nc, nc, nc)),
                             ("ldiph_lc", nc ** 3, (nc, nc, nc))):
        assert np.array_equal(got[key], model[o:o + size].reshape(shape)), key
        o += size
    assert got["rs_n"].max() > 1 and (got["ldiph_lc"] != dict2pid.BAD_SSID).any()


def test_triphone_lookup_backs_off_like_the_reference():
    md = s3io.read_mdef(os.path.join(EN, "mdef"))
    tri = dict2pid.TriphoneIndex(md)
    ci = {n: i for i, n in enumerate(md["ciname"])}
    n = md["n_ciphone"]
    assert tri.phone_id(ci["AA"], -1, ci["B"], 0) == ci["AA"] and tri.nearest(ci["AA"], ci["B"], -1, 1) == ci["AA"]
    hits = sum(tri.phone_id(ci["AA"], l, r, 0) >= n for l in range(n) for r in range(n))
    assert 0 < hits < n * n                                      # some word-internal triphones exist, not all
    for l in range(n):
        for r in range(n):
            p = tri.nearest(ci["AA"], l, r, dict2pid.WPOS_END)
            assert 0 <= p < md["n_phone"]
            if p < n:
                assert p == ci["AA"]                             # the last resort is the base phone itself
    noise = [i for i in range(n) if md["phone_filler"][i] and i != md["sil"]]
    assert noise and tri.phone_id(ci["AA"], noise[0], ci["B"], 0) == tri.phone_id(ci["AA"], md["sil"], ci["B"], 0)   # fillers count as silence


@pytest.fixture(scope="module")
def cmu():
    hd, dic = EN, os.path.join(REF, "model", "cmudict-en-us.dict")
    md = s3io.read_mdef(os.path.join(hd, "mdef"))
    words, prons, base, fs = lmio.read_dict(dic, os.path.join(hd, "noisedict"), md["ciname"])
    ci = {n: i for i, n in enumerate(md["ciname"])}
    pr = [[ci[x] for x in p] for p in prons]
    return md, {w: i for i, w in enumerate(words)}, pr, dict2pid.build(md, pr)


def test_alignment_phone_chains_equal_the_golden_ones(cmu):
    """ps_alignment_populate from the files alone: the (ssid, tmatid) chains tests/golden/en_us_align.npz holds for
    three transcripts -- the inputs the GPU-verified forced-alignment kernel is tested with."""
    from conftest import golden
    md, idx, pr, tabs = cmu
    g = golden("en_us_align.npz")
    for tag, text in (("a", "<s> go forward ten meters </s>"), ("b", "go forward ten meters"),
                      ("c", "<s> go forward ten meters </s> <s> go forward </s>")):
        s, t, c = dict2pid.alignment_phones(md, pr, tabs, [idx[w] for w in text.split()])
        assert np.array_equal(s, g[tag + "_ssid"]) and np.array_equal(t, g[tag + "_tmatid"]), tag


@pytest.mark.parametrize("text", ["a", "<s> a i </s>", "<sil> the a an </s>", "go <sil> forward(2) a ten", "++noise++ meters ++breath++ a"])
def test_alignment_phone_chains_live(cmu, text):
    """Single-phone words next to each other, fillers as neighbours, alternate pronunciations: against the reference."""
    md, idx, pr, tabs = cmu
    text = " ".join(w for w in text.split() if w in idx)
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    want = refdrv.align(EN, os.path.join(REF, "model", "cmudict-en-us.dict"), text, pcm)
    s, t, c = dict2pid.alignment_phones(md, pr, tabs, [idx[w] for w in text.split()])
    assert np.array_equal(s, want["ssid"]) and np.array_equal(t, want["tmatid"])


def test_keyphrase_chains_equal_the_golden_ones(cmu):
    """kws_search_reinit's keyphrase HMM chains from the files alone: the configuration tests/golden/en_us_kws.npz
    holds (what the GPU-verified keyword-spotting kernel is tested with), the phone-loop part included."""
    from conftest import golden
    md, idx, pr, tabs = cmu
    g = golden("en_us_kws.npz")
    n = md["n_ciphone"]
    # the key list "forward / ten meters / go / backward": the reference keeps it in reverse file order (glist prepend)
    for tag, phrases in (("a", ["forward"]), ("b", ["backward", "go", "ten meters", "forward"])):
        chains = [dict2pid.keyphrase_phones(md, pr, tabs, [idx[w] for w in ph.split()]) for ph in phrases]
        assert np.array_equal(np.concatenate([c[0] for c in chains]), g[tag + "_kp_ssid"])
        assert np.array_equal(np.concatenate([c[1] for c in chains]), g[tag + "_kp_tmat"])
        assert np.array_equal(np.cumsum([0] + [len(c[0]) for c in chains]), g[tag + "_kp_off"])
        assert np.array_equal(md["phone_ssid"][:n], g[tag + "_pl_ssid"]) and np.array_equal(md["phone_tmat"][:n], g[tag + "_pl_tmat"])
