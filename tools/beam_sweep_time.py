"""Device time of the search-scale Viterbi sweep with and without beam pruning between frames at the headline's shape
(bench.py: 1000 utterances x 998 frames, 6 081 entered instances per utterance, BASELINE senone count): the unpruned
fused sweep, the cluster sweep with a beam so wide that nothing leaves (cost of the per-frame cluster exchange alone),
and with a beam / -maxhmmpf that prune (work falls with the active set)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                    # noqa: E402
from pocketsphinx_b200 import api               # noqa: E402


def main():
    U, T = int(os.environ.get("UTTS", "1000")), 998
    pm, _desc, _raw = bench.load_model("baseline")
    ctx = api.HmmContext(pm.tp, pm.sseq, pm.n_sen, device=0)
    tmpl = bench.channel_template(pm, api.HMM_DTYPE)
    tmpl["frame"] = 0                        # active in frame 0
    rng = np.random.default_rng(3)
    tmpl["score"][:, 0] = -rng.integers(0, 20000, len(tmpl)).astype(np.int32)   # entry scores spread over the beam
    hs = api.HmmSet(ctx, U * bench.N_ACTIVE + U * 512, U)
    hs.upload(np.tile(tmpl, U), np.arange(U + 1, dtype=np.int64) * bench.N_ACTIVE)
    hs.snapshot()
    g = torch.Generator(device="cuda").manual_seed(1)
    scr = torch.randint(0, 900, (U * T, pm.n_sen), dtype=torch.int16, device="cuda", generator=g)
    d_row0 = torch.arange(U, dtype=torch.int64, device="cuda") * T
    d_best = torch.empty((T, U), dtype=torch.int32, device="cuda")
    d_nact = torch.empty((T, U), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    out = {"utts": U, "frames": T, "instances_per_utt": bench.N_ACTIVE, "n_sen": int(pm.n_sen)}

    only = os.environ.get("ONLY")

    def run(name, fn):
        if only and name != only:
            out[name] = {"ms": None}
            return out[name]
        ms = []
        for _ in range(int(os.environ.get("REPS", "3"))):
            hs.restore()
            ms.append(fn())
        out[name] = {"ms": min(ms)}
        return out[name]

    run("fused_sweep_no_pruning", lambda: hs.sweep_device(scr.data_ptr(), U * T, T, d_best.data_ptr(), d_row0=d_row0.data_ptr()))
    for name, beam, mh in [("cluster_sweep_beam_neutral", -0x1fffffff, -1), ("cluster_sweep_beam_1e4", -10000, -1),
                           ("cluster_sweep_beam_3000", -3000, -1), ("cluster_sweep_beam_1e4_maxhmmpf_3000", -10000, 3000)]:
        r = run(name, lambda: hs.sweep_beam_device(scr.data_ptr(), U * T, T, 0, beam, d_best.data_ptr(), maxhmmpf=mh,
                                                  d_n_active=d_nact.data_ptr(), d_row0=d_row0.data_ptr()))
        if r["ms"] is None:
            continue
        na = d_nact.cpu().numpy()
        r.update(beam=beam, maxhmmpf=mh, active_first=int(na[0].mean()), active_mean=float(na.mean()), active_last=int(na[-1].mean()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
