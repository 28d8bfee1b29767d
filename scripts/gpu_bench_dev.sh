#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/bench_dev.log
echo "=== unet tests" >> gpurun_out/bench_dev.log
timeout 1200 python -m pytest tests/test_unet_gpu.py -q -s -k "not svd_config" --no-header -p no:cacheprovider 2>&1 | tail -40 >> gpurun_out/bench_dev.log
echo "=== kernels" >> gpurun_out/bench_dev.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tapgemm_gpu.py -q --no-header -p no:cacheprovider 2>&1 | tail -30 >> gpurun_out/bench_dev.log
echo "=== bench eager" >> gpurun_out/bench_dev.log
SVDX_GEMM_TABLE=gpurun_out/gemm_table.json timeout 900 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.json 2>> gpurun_out/bench_dev.log
cat gpurun_out/bench_eager.json >> gpurun_out/bench_dev.log
echo "=== bench graph" >> gpurun_out/bench_dev.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.json 2>> gpurun_out/bench_dev.log
cat gpurun_out/bench_graph.json >> gpurun_out/bench_dev.log
echo "=== ncu launch list" >> gpurun_out/bench_dev.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --profile-one --warmup 1 --no-graph >> gpurun_out/bench_dev.log 2>&1
tail -c 4000 gpurun_out/bench_dev.log
