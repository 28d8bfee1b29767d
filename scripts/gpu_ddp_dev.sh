#!/bin/bash
# N=2 data-parallel bench (run with gpurun --gpus 2)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/ddp_dev.log 2>&1
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2>> gpurun_out/ddp_dev.log
cat gpurun_out/bench_n2.json >> gpurun_out/ddp_dev.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/ddp_check.py >> gpurun_out/ddp_dev.log 2>&1
tail -c 3000 gpurun_out/ddp_dev.log
