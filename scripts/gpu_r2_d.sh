#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2d.log
: > $L
echo "=== pytest -m gpu" >> $L
timeout 1800 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | cut -c1-300 | tail -60 >> $L
echo "=== kbench" >> $L
timeout 600 python scripts/kbench.py gn attn > gpurun_out/kbench_r2d.txt 2>&1; cat gpurun_out/kbench_r2d.txt >> $L
echo "--- SVDX_ATTN_SMALL=0" >> $L
SVDX_ATTN_SMALL=0 timeout 300 python scripts/kbench.py attn 2>&1 | grep temporal >> $L
echo "=== bench" >> $L
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench_r2d.json 2>> $L
python - >> $L <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2d.json'))
print("ms/step", d['ms_per_step'], "value", d['value'], "e2e", d['e2e']['value'], "launches", d['gpu_launches'], "loss", d['config']['final_loss'])
print("families", {k:(round(v['ms_per_step'],2), round(v['frac'],3), v['launches_per_step']) for k,v in d['roofline_by_family'].items()})
print("roofline", round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms_per_step'],2))
print("script_path", d.get('script_path'))
print("vae_encode", d.get('vae_encode'))
print("gpu_eager_baseline", d.get('gpu_eager_baseline'))
PY
echo "=== ncu launch list" >> $L
SVDX_SHAPE_LOG=gpurun_out/shapes_r2d.json timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_r2d.csv python bench.py --profile-one --warmup 1 --no-graph >> $L 2>&1
tail -c 2500 $L | cut -c1-400
