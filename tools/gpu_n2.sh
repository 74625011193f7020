#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "n2 weak exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --batch-total 1000 > gpurun_out/r02_bench_n2_strong.json 2> gpurun_out/r02_bench_n2_strong.err; echo "n2 strong exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/r02_bench_n2_ref.json 2> gpurun_out/r02_bench_n2_ref.err; echo "n2 ref exit $?"
for f in n2 n2_strong n2_ref; do python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.4g" % d["value"], d.get("ms_per_step"), d.get("scaling"), "e2e %.4g" % d["e2e"]["value"], d["config"].get("parallelism"), d["config"].get("model_broadcast_ms"))
except Exception as e:
    print("$f unreadable", e); print(open("gpurun_out/r02_bench_$f.err").read()[-800:])
P
done
