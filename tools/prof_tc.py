#!/usr/bin/env python
"""Score a synthetic BASELINE-shape batch a few times (profiling target for ncu).
   python tools/prof_tc.py [utts] [secs] [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pocketsphinx_b200 import api  # noqa: E402
from pocketsphinx_b200.model import synth_feats, synth_ptm  # noqa: E402

U = int(sys.argv[1]) if len(sys.argv) > 1 else 200
secs = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
model_kind = sys.argv[4] if len(sys.argv) > 4 else "ptm"
T = secs * 100 - 2
if model_kind == "cont":
    from pocketsphinx_b200.model import synth_ms
    pm = synth_ms(seed=0, n_sen=5138, n_density=8, featlens=(39,), topn=4)
else:
    pm = synth_ptm(seed=0)
feats = synth_feats(pm, U, T, seed=1234).reshape(U * T, pm.sumlen)
off = api.Batch.offsets([T] * U)
m = api.Model(pm)
b = api.Batch(m, U, U * T)
d = torch.from_numpy(feats).cuda()
scr = torch.empty((U * T, pm.n_sen), dtype=torch.int16, device="cuda")
for i in range(reps):
    t0 = time.perf_counter()
    b.score_device(d.data_ptr(), off, scr.data_ptr())
    b.sync()
    print("rep", i, "%.2f ms" % ((time.perf_counter() - t0) * 1e3), b.last_kernel_ms())
b.close(); m.close()
