#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2c.log
: > $L
echo "=== pytest kernels + tapgemm + unet + boundary" >> $L
timeout 1800 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | tail -40 >> $L
echo "=== kbench" >> $L
for c in 2 4 8; do echo "--- SVDX_GN_CTAS_PER_SM=$c" >> $L; SVDX_GN_CTAS_PER_SM=$c timeout 300 python scripts/kbench.py gn >> $L 2>&1; done
timeout 600 python scripts/kbench.py ln attn wgrad >> $L 2>&1
echo "=== bench" >> $L
timeout 1200 python bench.py --no-cpu-baseline --no-gpu-baseline --no-script-path > gpurun_out/bench_r2c.json 2>> $L
cat gpurun_out/bench_r2c.json >> $L
tail -c 3000 $L
