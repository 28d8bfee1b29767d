#!/bin/bash
# (historical A/B: the register-array GroupNorm / LayerNorm kernels selected by SVDX_GN_RING=0 / SVDX_LN_RING=0 / SVDX_LIB alt builds were removed after this comparison)
mkdir -p gpurun_out
L=gpurun_out/r2g.log
: > $L
echo "=== pytest (gemv / lora / norm kernels)" >> $L
timeout 900 python -m pytest tests/test_tapgemm_gpu.py tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_boundary_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "gemv or lora or groupnorm or layernorm or half_precision or residual" 2>&1 | grep -v "^$" | cut -c1-300 | tail -15 >> $L
for v in base gn2 gn2r2 gn1r8 gn2r8; do
  echo "--- variant $v" >> $L
  if [ $v == base ]; then timeout 200 python scripts/kbench.py gn 2>&1 | grep "M=" >> $L; else SVDX_LIB=svd_xtend_b200/lib/alt_$v/libsvdx_b200.so timeout 200 python scripts/kbench.py gn 2>&1 | grep "M=" >> $L; fi
done
for v in base ln3 ln4; do
  echo "--- variant $v" >> $L
  if [ $v == base ]; then timeout 200 python scripts/kbench.py ln 2>&1 | grep "M=" >> $L; else SVDX_LIB=svd_xtend_b200/lib/alt_$v/libsvdx_b200.so timeout 200 python scripts/kbench.py ln 2>&1 | grep "M=" >> $L; fi
done
echo "=== bench config 5" >> $L
timeout 1200 python bench.py --config 5 --no-cpu-baseline --no-gpu-baseline --no-script-path --no-families > gpurun_out/bench_r2g_c5.json 2>> $L
python - >> $L <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2g_c5.json').read().splitlines() if l.startswith('{')][-1])
print("config 5: ms/step", round(d['ms_per_step'],3), "value", round(d['value'],1), "launches", d['gpu_launches'], "loss", d['config']['final_loss'])
PY
grep -v "UserWarning\|frombuffer" $L | cut -c1-330 | tail -60
