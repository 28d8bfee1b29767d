#!/bin/bash
# N-GPU data-parallel validation (gpurun --gpus N -- bash scripts/gpu_ddp_n.sh N): equivalence script, then the bench with the
# peer-memory exchange and with the NCCL sharded exchange
N=${1:-4}
mkdir -p gpurun_out
L=gpurun_out/ddp_n$N.log
nvidia-smi -L > $L 2>&1
echo "=== ddp_check N=$N" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 scripts/ddp_check.py > gpurun_out/ddp_check_n$N.txt 2>&1
grep "ddp_check\]" gpurun_out/ddp_check_n$N.txt >> $L
for mode in p2p sharded; do
  echo "=== bench N=$N --ddp $mode" >> $L
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --ddp $mode --no-families --no-script-path --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_n${N}_$mode.json 2>> gpurun_out/ddp_n${N}_err.log
  python - >> $L <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n${N}_$mode.json').read().splitlines() if l.startswith('{')][-1])
    print("$mode", "ms/step", round(d['ms_per_step'],3), "value", round(d['value'],1), "parallelism", d['config']['parallelism'], "loss", d['config']['final_loss'], "exchange", d.get('exchange'))
except Exception as e:
    print("$mode failed", e)
PY
done
tail -c 3000 gpurun_out/ddp_n${N}_err.log | grep -v "UserWarning\|frombuffer\|OMP_NUM\|^\*\*\*\|^$" | tail -15 >> $L
cat $L | cut -c1-500
