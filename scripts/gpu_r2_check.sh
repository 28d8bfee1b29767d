#!/bin/bash
# round-2 check run: parity tests, smoke, default bench line (families + baselines), ncu launch list of one step
R=${1:-r2a}
mkdir -p gpurun_out
L=gpurun_out/check_$R.log
: > $L
echo "=== pytest -m gpu" >> $L
timeout 2400 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider -x -s 2>&1 | grep -v "^$" | tail -40 >> $L
echo "=== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
echo "=== bench (default)" >> $L
timeout 1500 python bench.py > gpurun_out/bench_$R.json 2>> $L
cat gpurun_out/bench_$R.json >> $L
if [ "$2" != "nolist" ]; then
echo "=== ncu launch list" >> $L
SVDX_SHAPE_LOG=gpurun_out/shapes_$R.json timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_$R.csv python bench.py --profile-one --warmup 1 --no-graph >> $L 2>&1
fi
tail -c 6000 $L
