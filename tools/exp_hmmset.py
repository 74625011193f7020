"""Scratch experiment (GPU): Viterbi-stage throughput of psb_hmmset_eval_frames_device alone
(random senone scores, 1000 utterances x 6081 instances, 3-state)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pocketsphinx_b200 import api
from pocketsphinx_b200.model import synth_ptm
pm = synth_ptm(seed=0)
U = int(os.environ.get("UTTS", "1000")); A = int(os.environ.get("ACTIVE", "6081")); F = 48
ns = pm.n_emit_state
ctx = api.HmmContext(pm.tp, pm.sseq, pm.n_sen)
rng = np.random.default_rng(99)
n = U * A
hm = np.zeros(n, api.HMM_DTYPE)
ssid = rng.integers(0, len(pm.sseq), n)
hm["score"][:, :] = -0x20000000; hm["score"][:, 0] = 0
hm["history"][:, :] = -1; hm["out_score"] = -0x20000000; hm["out_history"] = -1; hm["bestscore"] = -0x20000000
hm["ssid"] = ssid; hm["senid"][:, :ns] = pm.sseq[ssid]; hm["tmatid"] = rng.integers(0, pm.tp.shape[0], n); hm["n_emit_state"] = ns
hs = api.HmmSet(ctx, n + (U * 512 if os.environ.get('ALIGN_TILES', '1') == '1' else 0), U)
hs.upload(hm, np.arange(U + 1, dtype=np.int64) * A)
scr = torch.randint(0, 600, (F * U, pm.n_sen), dtype=torch.int16, device="cuda")
row0 = torch.arange(U, dtype=torch.int64, device="cuda") * F
best = torch.empty((F, U), dtype=torch.int32, device="cuda")
hs.eval_frames_device(scr.data_ptr(), 3, best.data_ptr(), d_row0=row0.data_ptr())
for rep in range(3):
    ms = hs.eval_frames_device(scr.data_ptr(), F, best.data_ptr(), d_row0=row0.data_ptr())
    alg = (2 * ns * 4) * 2 + 2 * 4 * 2 + 4 + 2 * ns + 2 + 2 * ns
    print("threads", os.environ.get("PSB_HMMSET_THREADS", "256"), "us/frame %.1f" % (ms / F * 1e3), "GB/s %.0f" % (n * alg * F / (ms * 1e-3) / 1e9), flush=True)
