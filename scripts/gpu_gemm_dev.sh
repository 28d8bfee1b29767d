#!/bin/bash
# bring-up run for the tcgen05 tapgemm kernel: each test group in its own process so that a trap
# in one does not poison the CUDA context of the others.

mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for k in "weight_gradient" "dot_diff" "linear_bias" "f32_out" "geglu" "residual_blend" "rowbias" "temporal_conv" "conv3x3 and not stride2" "stride2" "wgrad"; do
  echo "=== $k" >> gpurun_out/gemm_dev.log
  timeout 300 python -m pytest tests/test_tapgemm_gpu.py -q -k "$k" -x --no-header -p no:cacheprovider >> gpurun_out/gemm_dev.log 2>&1
  echo "exit $?" >> gpurun_out/gemm_dev.log
done
tail -c 6000 gpurun_out/gemm_dev.log
