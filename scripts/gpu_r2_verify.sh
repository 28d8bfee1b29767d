#!/bin/bash
# quick verification of the tree: all GPU tests, smoke, a short default-config bench
mkdir -p gpurun_out
L=gpurun_out/verify.log
: > $L
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | cut -c1-300 | tail -6 >> $L
timeout 600 python __graft_entry__.py smoke 2>&1 | grep -v Warning | tail -2 >> $L
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-script-path --no-families > gpurun_out/bench_verify.json 2>> gpurun_out/verify_err.log
python - >> $L <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_verify.json').read().splitlines() if l.startswith('{')][-1])
print("bench: ms/step", round(d['ms_per_step'],3), "value", round(d['value'],1), "launches", d['gpu_launches'], "loss", d['config']['final_loss'])
PY
grep -v "UserWarning\|frombuffer" $L | cut -c1-300
