#!/bin/bash
# end-of-round artefacts on a short GPU budget: default bench line, ncu launch list, per-launch dram bytes, one full capture
R=${1:-r1}
mkdir -p gpurun_out
: > gpurun_out/final.log
echo "=== bench (default)" >> gpurun_out/final.log
timeout 400 python bench.py > gpurun_out/bench_default.json 2>> gpurun_out/final.log
cat gpurun_out/bench_default.json >> gpurun_out/final.log
echo "=== ncu launch list" >> gpurun_out/final.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_$R.csv python bench.py --profile-one --warmup 1 --no-graph >> gpurun_out/final.log 2>&1
echo "=== ncu dram bytes of every tapgemm launch" >> gpurun_out/final.log
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tapgemm -c 4000 --csv --log-file gpurun_out/tapgemm_dram_$R.csv python bench.py --profile-one --warmup 1 --no-graph >> gpurun_out/final.log 2>&1
echo "=== ncu full tapgemm" >> gpurun_out/final.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tapgemm2 -s 2600 -c 3 -f -o gpurun_out/prof_tapgemm_$R python bench.py --profile-one --warmup 1 --no-graph >> gpurun_out/final.log 2>&1
ls -la gpurun_out/*.ncu-rep >> gpurun_out/final.log 2>&1
tail -c 2500 gpurun_out/final.log
