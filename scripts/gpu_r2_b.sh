#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2b.log
: > $L
echo "=== graph gap probe" >> $L
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/graph_gap_probe scripts/probes/graph_gap_probe.cu >> $L 2>&1 && timeout 300 gpurun_out/graph_gap_probe >> $L 2>&1
echo "=== determinism probe" >> $L
timeout 300 python scripts/determinism_probe.py >> $L 2>&1
echo "=== pytest -m gpu (no -x)" >> $L
timeout 2400 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | tail -60 >> $L
tail -c 5000 $L
