import os, sys, time, subprocess
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pocketsphinx_b200 import api
from pocketsphinx_b200.model import synth_ptm, synth_feats
import bench
pm = synth_ptm(seed=0)
T=998; U=1000
feats = synth_feats(pm, U, T, seed=1234)
model = api.Model(pm)
total=U*T
off = api.Batch.offsets([T]*U)
fp = torch.from_numpy(feats.reshape(total, pm.sumlen)).pin_memory()
d = fp.cuda()
batch = api.Batch(model, U, total)
ctx = api.HmmContext(pm.tp, pm.sseq, pm.n_sen)
H = pm.n_ciphone
PL = bench.PL
pl = api.PhoneLoop(ctx, pm.phone_ssid[:H], pm.phone_tmat[:H], PL["window"], PL["beam"], PL["pbeam"], PL["pip"], PL["weight"])
def run(n):
    batch.event_record(0)
    for _ in range(n): batch.decode_device(pl, d.data_ptr(), off)
    batch.event_record(1)
    return batch.event_elapsed_ms()/n
for _ in range(3): batch.decode_device(pl, d.data_ptr(), off)
batch.sync()
print("no sampler", run(5), run(5))
for lms in ("100","200","500"):
    p = subprocess.Popen(["nvidia-smi","-i","0","--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap,power.draw","--format=csv,noheader,nounits","-lms",lms], stdout=subprocess.PIPE, text=True)
    time.sleep(0.5)
    print("sampler", lms, run(5), run(5))
    p.terminate(); out=p.stdout.read().strip().splitlines(); print(len(out), out[-3:])
print("no sampler", run(5), run(5))
