#!/bin/bash
# bring-up run: attention / norm / elementwise kernels, one process per group
mkdir -p gpurun_out
: > gpurun_out/kernels_dev.log
for k in "attention_spatial" "attention_temporal" "groupnorm" "layernorm" "prep_weight" "layout" "misc"; do
  echo "=== $k" >> gpurun_out/kernels_dev.log
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "$k" --no-header -p no:cacheprovider 2>&1 | tail -40 >> gpurun_out/kernels_dev.log
  echo "exit $?" >> gpurun_out/kernels_dev.log
done
tail -c 8000 gpurun_out/kernels_dev.log
