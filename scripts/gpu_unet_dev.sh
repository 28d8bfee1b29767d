#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/unet_dev.log
for k in "tiny_forward_backward" "tiny_inference" "svd_config_forward"; do
  echo "=== $k" >> gpurun_out/unet_dev.log
  timeout 900 python -m pytest tests/test_unet_gpu.py -q -s -k "$k" --no-header -p no:cacheprovider 2>&1 | tail -60 >> gpurun_out/unet_dev.log
  echo "exit $?" >> gpurun_out/unet_dev.log
done
tail -c 10000 gpurun_out/unet_dev.log
