"""CPU: the C restatement of fsg_search.c + fsg_history.c (oracle/ps_oracle.c: pso_fsg_run) against
what the reference's own fsg_search produced on goforward.raw (tests/golden/en_us_fsg.npz): every
history-table entry (link, frame, score, predecessor, left context, right-context bit vector), the
exit the reference picks and its hypothesis, on the reference's senone scores."""
import numpy as np
import pytest

from conftest import golden

TAGS = ("go", "go_hmmpf", "cmd", "cmd_wide", "cmd_hmmpf")


def _case(g, tag):
    return {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + ".")}


@pytest.mark.parametrize("tag", TAGS)
def test_fsg_oracle_matches_reference_golden(tag):
    from oracle import oracle
    m = golden("en_us_ptm_model.npz")
    scr = golden("en_us_goforward.npz")["senscr"]
    c = _case(golden("en_us_fsg.npz"), tag)
    hist = oracle.fsg_run(m["tp"], m["sseq"], c, scr)
    assert hist.shape == c["hist"].shape and np.array_equal(hist, c["hist"])
    bp, score = oracle.fsg_find_exit(hist, c["links"], len(scr), int(c["final_state"]))
    assert bp > 0 and score == int(c["score"])
    vocab = str(c["vocab"]).split("\n")
    words = [vocab[w] for w in oracle.fsg_hyp_wids(hist, c["links"], bp)]
    # fsg_search_hyp leaves fillers out of the string (fsg_search.c:1040-1050: dict_real_word)
    assert " ".join(w for w in words if not w.startswith("<") and not w.startswith("+")) == str(c["hyp"])


def test_fsg_golden_covers_the_interesting_paths():
    g = golden("en_us_fsg.npz")
    go, hp, cmd = _case(g, "go"), _case(g, "go_hmmpf"), _case(g, "cmd")
    assert len(hp["hist"]) < len(go["hist"])                     # -maxhmmpf narrowed the beams
    assert (cmd["hist"][:, 0] >= 0).sum() and len(cmd["nullarc"]) > len(go["nullarc"])
    null_links = set(np.nonzero(cmd["links"][:, 2] < 0)[0].tolist())
    assert any(int(l) in null_links for l in cmd["hist"][:, 0])  # null transitions were taken
    assert cmd["links"][:, 4].any()                              # fillers / single-phone words: all right contexts
    partial = (cmd["hist"][:, 5:] != -1).any(1) & (cmd["hist"][:, 5:] != 0).any(1)
    assert partial.any()                                         # right-context subtraction left partial sets
