"""GPU (-m gpu): psb_ngram_fwdtree_batch_device (the first pass of the n-gram search on the device)
against the reference's golden backpointer tables and against the oracle on ragged batches.

STATUS: ngs_fwdtree_kernel was written after this round's GPU minutes were spent.  Its phase code
is checked on the host against the reference (tests/test_ngs_emul.py), the kernel itself has not
run on hardware yet: this file only runs with PSB_RUN_UNVERIFIED=1.  Nothing in DESIGN.md claims
device parity for this path."""
import os

import numpy as np
import pytest

from conftest import golden

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PSB_RUN_UNVERIFIED") != "1",
                                 reason="ngs_fwdtree_kernel not yet run on hardware (set PSB_RUN_UNVERIFIED=1)")]

TAGS = ("default", "wide", "narrow", "maxwpf", "abs", "pen", "lookahead")


def _case(g, tag):
    return {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + ".")}


@pytest.fixture(scope="module")
def api():
    from pocketsphinx_b200 import api
    assert api.device_count() > 0, "no CUDA device visible"
    return api


@pytest.mark.timeout(300)
@pytest.mark.parametrize("tag", TAGS)
def test_fwdtree_batch_matches_reference_and_oracle(api, en_us, tag):
    import torch
    from oracle import oracle
    gf = golden("en_us_goforward.npz")
    scr = gf["senscr"]
    c = _case(golden("en_us_fwdtree.npz"), tag)
    n_ci = int(c["info"][6])
    cit = en_us.phone_tmat[:n_ci]
    parts = [scr, scr[:0], scr[:1], scr[:100], scr]
    utt_off = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int32)
    d_scr = torch.from_numpy(np.ascontiguousarray(np.concatenate(parts))).cuda()
    d_pen, pens = None, [None] * len(parts)
    if tag == "lookahead":
        win = int(gf["pl_params"][4])
        # penalties in force per search frame; prefixes of the utterance see the same phone-loop history
        pens = [np.ascontiguousarray(gf["pl_pen"][np.minimum(np.arange(len(p)) + win, max(len(p) - 1, 0))], np.int32)
                if len(p) else np.zeros((0, n_ci), np.int32) for p in parts]
        d_pen = torch.from_numpy(np.concatenate(pens)).cuda()
    ctx = api.HmmContext(en_us.tp, en_us.sseq, en_us.n_sen)
    out = ctx.ngram_fwdtree(d_scr.data_ptr(), utt_off, c["info"], c["model"], cit, len(c["bp"]) + 64, len(c["bss"]) + 4096,
                            d_pen.data_ptr() if d_pen is not None else None)
    for u in (0, 4):                                            # the reference's own tables
        bp, bss, idx = out[u]
        assert np.array_equal(bp, c["bp"]) and np.array_equal(bss, c["bss"]) and np.array_equal(idx, c["bp_idx"]), u
    for u in (1, 2, 3):
        want = oracle.fwdtree_run(en_us.tp, en_us.sseq, cit, c["info"], c["model"], parts[u], pen_in_force=pens[u])
        bp, bss, idx = out[u]
        assert np.array_equal(bp, want[0]) and np.array_equal(bss, want[1]) and np.array_equal(idx, want[2]), u
    ctx.close()


@pytest.mark.timeout(300)
def test_fwdtree_full_table_is_an_error(api, en_us):
    import torch
    from pocketsphinx_b200._lib import PsbError
    gf = golden("en_us_goforward.npz")
    c = _case(golden("en_us_fwdtree.npz"), "default")
    d_scr = torch.from_numpy(np.ascontiguousarray(gf["senscr"])).cuda()
    ctx = api.HmmContext(en_us.tp, en_us.sseq, en_us.n_sen)
    with pytest.raises(PsbError):
        ctx.ngram_fwdtree(d_scr.data_ptr(), np.array([0, 278], np.int32), c["info"], c["model"],
                          en_us.phone_tmat[:int(c["info"][6])], 100, 100000)
    ctx.close()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("tag", ("flat_default", "flat_wide", "flat_narrow"))
def test_fwdflat_batch_matches_reference_and_oracle(api, en_us, tag):
    """Second pass on the device behind the first pass on the device (ngs_fwdflat_kernel)."""
    import torch
    from oracle import oracle
    gf = golden("en_us_goforward.npz")
    scr = gf["senscr"]
    c = _case(golden("en_us_fwdtree.npz"), tag)
    n_ci = int(c["info"][6])
    cit, cis = en_us.phone_tmat[:n_ci], en_us.phone_ssid[:n_ci]
    parts = [scr, scr[:0], scr[:1], scr[:120], scr]
    utt_off = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int32)
    d_scr = torch.from_numpy(np.ascontiguousarray(np.concatenate(parts))).cuda()
    d_pen, pens = None, [None] * len(parts)
    if tag == "flat_default":
        win = int(gf["pl_params"][4])
        pens = [np.ascontiguousarray(gf["pl_pen"][np.minimum(np.arange(len(p)) + win, max(len(p) - 1, 0))], np.int32)
                if len(p) else np.zeros((0, n_ci), np.int32) for p in parts]
        d_pen = torch.from_numpy(np.concatenate(pens)).cuda()
    ctx = api.HmmContext(en_us.tp, en_us.sseq, en_us.n_sen)
    first = ctx.ngram_fwdtree(d_scr.data_ptr(), utt_off, c["info"], c["model"], cit, 8192, 1 << 18,
                              d_pen.data_ptr() if d_pen is not None else None)
    out = ctx.ngram_fwdflat(d_scr.data_ptr(), utt_off, c["info"], c["model"], cit, cis, [t[0] for t in first],
                            len(c["bp"]) + 64, len(c["bss"]) + 4096)
    for u in (0, 4):
        bp, bss, idx = out[u]
        assert np.array_equal(bp, c["bp"]) and np.array_equal(bss, c["bss"]) and np.array_equal(idx, c["bp_idx"]), u
    for u in (1, 2, 3):
        bp1 = oracle.fwdtree_run(en_us.tp, en_us.sseq, cit, c["info"], c["model"], parts[u], pen_in_force=pens[u])[0]
        assert np.array_equal(first[u][0], bp1), u
        want = oracle.fwdflat_run(en_us.tp, en_us.sseq, cit, cis, c["info"], c["model"], bp1, parts[u])
        bp, bss, idx = out[u]
        assert np.array_equal(bp, want[0]) and np.array_equal(bss, want[1]) and np.array_equal(idx, want[2]), u
    ctx.close()
