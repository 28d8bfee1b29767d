#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/bench_dev.log
for k in "tiny_inference" "checkpointing" "lora" "full_finetune" "tiny_forward_backward" "svd_config_forward"; do
  echo "=== $k" >> gpurun_out/bench_dev.log
  timeout 900 python -m pytest tests/test_unet_gpu.py -q -s -k "$k" --no-header -p no:cacheprovider 2>&1 | tail -30 >> gpurun_out/bench_dev.log
done
echo "=== kernels" >> gpurun_out/bench_dev.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tapgemm_gpu.py -q --no-header -p no:cacheprovider 2>&1 | tail -30 >> gpurun_out/bench_dev.log
echo "=== graph debug" >> gpurun_out/bench_dev.log
timeout 900 python scripts/graph_debug.py --full >> gpurun_out/bench_dev.log 2>&1
echo "=== bench eager" >> gpurun_out/bench_dev.log
SVDX_GEMM_TABLE=gpurun_out/gemm_table.json timeout 900 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.json 2>> gpurun_out/bench_dev.log
cat gpurun_out/bench_eager.json >> gpurun_out/bench_dev.log
echo "=== ncu launch list" >> gpurun_out/bench_dev.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --profile-one --warmup 1 --no-graph >> gpurun_out/bench_dev.log 2>&1
tail -c 6000 gpurun_out/bench_dev.log
