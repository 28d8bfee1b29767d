#!/bin/bash
# bring-up of the CTA-pair (tcgen05 cta_group::2) tapgemm kernel
mkdir -p gpurun_out
: > gpurun_out/pair_dev.log
export SVDX_2CTA=1
for k in "linear_bias" "geglu" "residual_blend or rowbias" "temporal_conv and not weight" "conv3x3 and not weight and not stride2 and not split" "stride2 and not weight"; do
  echo "=== $k" >> gpurun_out/pair_dev.log
  timeout 300 python -m pytest tests/test_tapgemm_gpu.py -q -x -k "$k" --no-header -p no:cacheprovider 2>&1 | tail -25 >> gpurun_out/pair_dev.log
done
echo "=== unet" >> gpurun_out/pair_dev.log
timeout 900 python -m pytest tests/test_unet_gpu.py -q -s -x -k "tiny_forward_backward or svd_config_forward" --no-header -p no:cacheprovider 2>&1 | tail -25 >> gpurun_out/pair_dev.log
echo "=== bench 2cta" >> gpurun_out/pair_dev.log
SVDX_GEMM_TABLE=gpurun_out/gemm_table_2cta.json timeout 900 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_2cta.json 2>> gpurun_out/pair_dev.log
cat gpurun_out/bench_2cta.json >> gpurun_out/pair_dev.log
tail -c 5000 gpurun_out/pair_dev.log
