"""Scratch experiment (GPU): kernel time of psb_align_batch_device, 1000 utterances x 998 frames,
100 phones each, random scores.  PSB_ALIGN_SEQ_SCAN=1 selects the one-thread transition scan."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pocketsphinx_b200 import api
from pocketsphinx_b200.model import synth_ptm
pm = synth_ptm(seed=0)
U, T, NP = 1000, 998, int(os.environ.get("PHONES", "100"))
ctx = api.HmmContext(pm.tp, pm.sseq, pm.n_sen)
rng = np.random.default_rng(5)
off = (np.arange(U + 1) * T).astype(np.int32)
ph_off = (np.arange(U + 1) * NP).astype(np.int32)
ssid = rng.integers(0, len(pm.sseq), U * NP).astype(np.int32)
tmat = rng.integers(0, pm.tp.shape[0], U * NP).astype(np.int32)
scr = torch.randint(0, 500, (U * T, pm.n_sen), dtype=torch.int16, device="cuda")
for rep in range(3):
    t0 = time.perf_counter()
    status, st, du, sc = ctx.align(None, off, ph_off, ssid, tmat, device_ptr=scr.data_ptr())
    dt = time.perf_counter() - t0
    print("seq" if os.environ.get("PSB_ALIGN_SEQ_SCAN") else "par", "phones", NP, "wall ms %.1f" % (dt * 1e3),
          "kernel ms %.2f" % api.lib().psb_align_last_kernel_ms(ctx.h), "ok", int((status == 0).sum()), flush=True)
