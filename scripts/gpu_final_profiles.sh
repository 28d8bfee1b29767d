#!/bin/bash
# round artefacts: parity tests, default bench line, ncu launch list + dram bytes + full captures (copied to profiles/ by scripts/summarize_profiles.py)
R=${1:-r1}
mkdir -p gpurun_out
: > gpurun_out/final.log
echo "=== pytest -m gpu" >> gpurun_out/final.log
timeout 1500 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -15 >> gpurun_out/final.log
echo "=== smoke" >> gpurun_out/final.log
timeout 300 python __graft_entry__.py smoke >> gpurun_out/final.log 2>&1
echo "=== bench (default)" >> gpurun_out/final.log
SVDX_GEMM_TABLE=gpurun_out/gemm_table.json timeout 1200 python bench.py > gpurun_out/bench_default.json 2>> gpurun_out/final.log
cat gpurun_out/bench_default.json >> gpurun_out/final.log
echo "=== ncu launch list" >> gpurun_out/final.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_$R.csv python bench.py --profile-one --warmup 1 --no-graph >> gpurun_out/final.log 2>&1
echo "=== ncu dram bytes of every tapgemm launch" >> gpurun_out/final.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tapgemm -c 4000 --csv --log-file gpurun_out/tapgemm_dram_$R.csv python bench.py --profile-one --warmup 1 --no-graph >> gpurun_out/final.log 2>&1
echo "=== ncu full tapgemm" >> gpurun_out/final.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tapgemm2 -s 2600 -c 5 -f -o gpurun_out/prof_tapgemm_$R python bench.py --profile-one --warmup 1 --no-graph >> gpurun_out/final.log 2>&1
echo "=== attention profile" >> gpurun_out/final.log
bash scripts/gpu_prof_attn.sh >> gpurun_out/final.log 2>&1
ls -la gpurun_out/*.ncu-rep >> gpurun_out/final.log 2>&1
tail -c 3000 gpurun_out/final.log
