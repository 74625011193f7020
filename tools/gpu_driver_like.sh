#!/bin/bash
# what the driver runs at round end, in its order
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_driver_tests.log 2>&1; echo "pytest -m gpu exit $?: $(tail -n 2 gpurun_out/r02_driver_tests.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_driver_smoke.log 2>&1; echo "smoke exit $?: $(tail -n 1 gpurun_out/r02_driver_smoke.log)"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_driver_ref.json 2> gpurun_out/r02_driver_ref.err; echo "reference arm exit $?: $(head -c 400 gpurun_out/r02_driver_ref.json)"
timeout 900 python bench.py > gpurun_out/r02_driver_bench.json 2> gpurun_out/r02_driver_bench.err; echo "bench exit $?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_driver_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "gpu_launches")}, d["e2e"]["value"], d["clocks"])
P
