#!/bin/bash
mkdir -p gpurun_out
PSB_TC_CHECK=1 timeout 180 python - > gpurun_out/r02_tc5_first.log 2>&1 <<'P'
import sys, numpy as np
sys.path.insert(0, ".")
from pocketsphinx_b200 import api
from pocketsphinx_b200.model import synth_feats, synth_ptm
from oracle import oracle
pm = synth_ptm(seed=0); f = synth_feats(pm, 6, 50, seed=5)
m = api.Model(pm); b = api.Batch(m, 8, 400)
off = api.Batch.offsets([50] * 6)
scr = b.score_host(f.reshape(-1, pm.sumlen), off)
om = oracle.OracleModel(pm)
ok = [bool(np.array_equal(scr[off[u]:off[u + 1]], om.score_utt(f[u]))) for u in range(6)]
print("identical", ok, "check", b.tc_check(), b.tc_stats)
P
echo "first run exit $?"; tail -5 gpurun_out/r02_tc5_first.log
