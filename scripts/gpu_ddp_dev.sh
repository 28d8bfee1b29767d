#!/bin/bash
# N=2 data-parallel checks (run with gpurun --gpus 2): equivalence script, then the bench in both exchange modes
mkdir -p gpurun_out
L=gpurun_out/ddp_dev.log
nvidia-smi -L > $L 2>&1
echo "=== kernel unit test (one device)" >> $L
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -k adamw_p2p 2>&1 | grep -v "^$" | tail -5 >> $L
echo "=== ddp_check" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/ddp_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -30 >> $L
for mode in p2p; do
  echo "=== bench N=2 --ddp $mode" >> $L
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --ddp $mode --no-families --no-script-path --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_n2_$mode.json 2>> $L
  python - >> $L <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n2_$mode.json'))
    print("$mode", "ms/step", round(d['ms_per_step'],3), "value", round(d['value'],1), "parallelism", d['config']['parallelism'], "loss", d['config']['final_loss'], "exchange", d.get('exchange'))
except Exception as e:
    print("$mode failed", e)
PY
done
echo "=== bench N=1 on the same box" >> $L
timeout 600 python bench.py --steps 10 --warmup 3 --no-families --no-script-path --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_n1_samebox.json 2>> $L
python - >> $L <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n1_samebox.json'))
    print("N=1", "ms/step", round(d['ms_per_step'],3), "value", round(d['value'],1))
except Exception as e:
    print("N=1 failed", e)
PY
grep -v "UserWarning\|frombuffer\|OMP_NUM_THREADS\|^\*\*\*\|^$" $L | cut -c1-400 | tail -40
