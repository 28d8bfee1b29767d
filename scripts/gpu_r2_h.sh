#!/bin/bash
# (historical A/B: the register-array GroupNorm / LayerNorm kernels selected by SVDX_GN_RING=0 / SVDX_LN_RING=0 / SVDX_LIB alt builds were removed after this comparison)
# cp.async ring norm kernels: correctness, A/B against the register-array kernels, CTAs-per-SM sweep; gemm block_n sweep; quick bench
mkdir -p gpurun_out
L=gpurun_out/r2h.log
: > $L
echo "=== pytest (norm kernels, ring on)" >> $L
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "groupnorm or layernorm or gn_ or ln_ or norm" 2>&1 | grep -v "^$" | cut -c1-300 | tail -15 >> $L
echo "--- gn base (register arrays)" >> $L
SVDX_GN_RING=0 timeout 200 python scripts/kbench.py gn 2>&1 | grep "M=" >> $L
for c in 2 3 4 6 8; do
  echo "--- gn ring cps=$c" >> $L
  SVDX_GN_RING_CPS=$c timeout 200 python scripts/kbench.py gn 2>&1 | grep "M=" >> $L
done
echo "--- ln base" >> $L
SVDX_LN_RING=0 timeout 200 python scripts/kbench.py ln 2>&1 | grep "M=" >> $L
for c in 1 2 3 4; do
  echo "--- ln ring fwd cps=$c" >> $L
  SVDX_LN_RING_CPS=$c timeout 200 python scripts/kbench.py ln 2>&1 | grep "M=" >> $L
done
echo "=== kbench gemm" >> $L
timeout 600 python scripts/kbench.py gemm 2>&1 | grep -v Warning | tail -60 >> $L
echo "=== full gpu tests" >> $L
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x 2>&1 | grep -v "^$" | cut -c1-300 | tail -8 >> $L
echo "=== bench config 2 (ring on)" >> $L
timeout 1200 python bench.py --no-cpu-baseline --no-gpu-baseline --no-script-path > gpurun_out/bench_r2h_c2.json 2>> $L
python - >> $L <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2h_c2.json').read().splitlines() if l.startswith('{')][-1])
print("config 2: ms/step", round(d['ms_per_step'],3), "value", round(d['value'],1), "launches", d['gpu_launches'], "loss", d['config']['final_loss'])
for k,v in d.get('roofline_by_family',{}).items(): print("  ", k, round(v['ms_per_step'],3), "ms  frac", round(v['frac'],3))
PY
grep -v "UserWarning\|frombuffer" $L | cut -c1-330 | tail -120
