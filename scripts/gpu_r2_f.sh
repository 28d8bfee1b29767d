#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2f.log
: > $L
echo "=== pytest -m gpu" >> $L
timeout 1800 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | cut -c1-400 | tail -60 >> $L
show() {
python - >> $L <<PY
import json
try:
    d=json.loads([l for l in open('$1').read().splitlines() if l.startswith('{')][-1])
    print("$1: ms/step", round(d['ms_per_step'],3), "value", round(d['value'],1), "e2e", round(d['e2e']['value'],1), "launches", d['gpu_launches'], "loss", d['config']['final_loss'], "workload", d['config']['workload'][:60])
    print("   families", {k:(round(v['ms_per_step'],2), round(v['frac'],3), v['launches_per_step']) for k,v in (d.get('roofline_by_family') or {}).items()})
    print("   roofline", round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms_per_step'],2), "TFLOP/step", round(d['roofline']['algorithmic_tflop_per_step'],2))
    sp=d.get('script_path') or {}
    print("   script_path", {k:(round(v,2) if isinstance(v,float) else v) for k,v in sp.items() if 'what' not in k})
    print("   vae", {k:v for k,v in (d.get('vae_encode') or {}).items() if k!='what'}, "gpu_eager", {k:v for k,v in (d.get('gpu_eager_baseline') or {}).items() if k!='what'})
except Exception as e:
    print("$1 failed", repr(e))
PY
}
echo "=== bench config 2" >> $L
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench_r2f_c2.json 2>> $L; show gpurun_out/bench_r2f_c2.json
echo "=== bench config 5 (LoRA r=64)" >> $L
timeout 1200 python bench.py --config 5 --no-cpu-baseline --no-gpu-baseline --no-script-path > gpurun_out/bench_r2f_c5.json 2>> $L; show gpurun_out/bench_r2f_c5.json
echo "=== bench config 4 (25 x 72 x 128, grad ckpt)" >> $L
timeout 1500 python bench.py --config 4 --steps 3 --no-cpu-baseline --no-gpu-baseline --no-script-path --no-families > gpurun_out/bench_r2f_c4.json 2>> $L; show gpurun_out/bench_r2f_c4.json
nvidia-smi --query-gpu=memory.used --format=csv >> $L
grep -v "UserWarning\|frombuffer" $L | tail -c 3500 | cut -c1-500
