#!/bin/bash
# ncu --set full of one launch of every hot kernel at its config-2 shape (scripts/prof_shapes.py) + the step launch list
R=${1:-r2}
mkdir -p gpurun_out
L=gpurun_out/prof_$R.log
: > $L
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'tapgemm|attn_|gn_|ln_|adamw|geglu|gemv' -f -o gpurun_out/prof_$R python scripts/prof_shapes.py >> $L 2>&1
ls -la gpurun_out/prof_$R.ncu-rep >> $L 2>&1
if [ "$2" == "list" ]; then
SVDX_SHAPE_LOG=gpurun_out/shapes_$R.json timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_$R.csv python bench.py --profile-one --warmup 1 --no-graph >> $L 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tapgemm -c 4000 --csv --log-file gpurun_out/tapgemm_dram_$R.csv python bench.py --profile-one --warmup 1 --no-graph >> $L 2>&1
fi
tail -c 1500 $L
