#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2e.log
: > $L
echo "=== pytest -m gpu" >> $L
timeout 1800 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | cut -c1-400 | tail -70 >> $L
echo "=== kbench" >> $L
timeout 600 python scripts/kbench.py gn attn > gpurun_out/kbench_r2e.txt 2>&1; cat gpurun_out/kbench_r2e.txt >> $L
for small in 1 0; do
echo "=== bench SVDX_ATTN_SMALL=$small" >> $L
SVDX_ATTN_SMALL=$small timeout 1500 python bench.py --no-cpu-baseline --no-gpu-baseline --no-script-path > gpurun_out/bench_r2e_$small.json 2>> $L
python - >> $L <<PY
import json
d=json.load(open('gpurun_out/bench_r2e_$small.json'))
print("ms/step", d['ms_per_step'], "value", d['value'], "e2e", d['e2e']['value'], "launches", d['gpu_launches'], "loss", d['config']['final_loss'])
print("families", {k:(round(v['ms_per_step'],2), round(v['frac'],3), v['launches_per_step']) for k,v in d['roofline_by_family'].items()})
print("roofline", round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms_per_step'],2))
PY
done
tail -c 2500 $L | cut -c1-400
