"""CPU, build container (needs oracle/_ref/libpsref.so): random grammars -- few states, many
homophones (two/to/too, for/four/fore, ...), null arcs, loops, random beams and -maxhmmpf -- decoded
by the reference's own fsg_search LIVE, against (1) the oracle restatement and (2) the host build of
the device search's phase code in both thread orders.  Every history-table row must agree.
PSB_FSG_FUZZ_SEEDS=a:b widens the range (175 seeds were run when this was written: 0 mismatches,
38 k equal-score exit ties)."""
import os
import random

import numpy as np
import pytest

from oracle import oracle, refdrv
from test_fsg_emul import _run, emul  # noqa: F401  (fixture)

pytestmark = pytest.mark.skipif(not refdrv.available(), reason="oracle/_ref/libpsref.so not built")
REF = os.path.dirname(refdrv.LIB_PATH)
WORDS = ("go forward backward ten two to too for four fore meters meter a i the eye aye one won right write left "
         "turn and an in inn no know oh owe eight ate be bee by buy bye see sea").split()


def _grammar(rng, path):
    ns = rng.randint(2, 7)
    lines = ["FSG_BEGIN fuzz", "NUM_STATES %d" % ns, "START_STATE 0", "FINAL_STATE %d" % (ns - 1)]
    seen = set()
    for _ in range(rng.randint(ns, 5 * ns)):
        a, b = rng.randrange(ns), rng.randrange(ns)
        w = None if rng.random() < 0.15 else rng.choice(WORDS)
        if (a, b, w) in seen or (w is None and a == b):
            continue
        seen.add((a, b, w))
        lines.append("TRANSITION %d %d %.3f%s" % (a, b, rng.uniform(0.05, 1), " " + w if w else ""))
    for s in range(ns - 1):                                     # a path from start to final always exists
        lines.append("TRANSITION %d %d 0.5 %s" % (s, s + 1, rng.choice(WORDS)))
    lines.append("FSG_END")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


@pytest.fixture(scope="module")
def scored():
    ref = refdrv.RefModel(os.path.join(REF, "model", "en-us"))
    pcm = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    pk = ref.packed()
    scr = np.ascontiguousarray(ref.score(ref.featurize_fresh(pcm)))
    ref.close()
    return pk, pcm, scr


def _seeds():
    a, b = (int(x) for x in os.environ.get("PSB_FSG_FUZZ_SEEDS", "0:6").split(":"))
    return list(range(a, b))


@pytest.mark.parametrize("seed", _seeds())
def test_random_grammar_reference_oracle_and_phase_code_agree(emul, scored, seed, tmp_path):  # noqa: F811
    pk, pcm, scr = scored
    rng = random.Random(seed)
    path = str(tmp_path / "fuzz.fsg")
    _grammar(rng, path)
    kv = {}
    if rng.random() < 0.3:
        kv.update(beam="1e-%d" % rng.randint(20, 90), pbeam="1e-%d" % rng.randint(20, 90), wbeam="1e-%d" % rng.randint(10, 70))
    if rng.random() < 0.3:
        kv["maxhmmpf"] = str(rng.randint(5, 200))
    r = refdrv.fsg(os.path.join(REF, "model", "en-us"), os.path.join(REF, "model", "cmudict-en-us.dict"), path, pcm, **kv)
    want = r["hist"]
    got = oracle.fsg_run(pk["tp"], pk["sseq"], r, scr)
    assert got.shape == want.shape and np.array_equal(got, want)
    hist, n = _run(emul, pk, r, scr, len(want) + 8)
    assert n == len(want) and np.array_equal(hist, want)
