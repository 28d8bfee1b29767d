#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2n.log
: > $L
echo "=== kbench gnb res (with the L2 prefetch of epilogue operands)" >> $L
timeout 600 python scripts/kbench.py gnb 2>&1 | grep -v Warning | tail -24 >> $L
echo "=== pytest gnb" >> $L
timeout 1500 python -m pytest tests/test_tapgemm_gpu.py -q -m gpu --no-header -p no:cacheprovider -k backward_sums 2>&1 | grep -v "^$" | cut -c1-400 | tail -6 >> $L
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-script-path --no-families"
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 $B > gpurun_out/bench_r2n_$tag.json 2>> gpurun_out/r2n_err.log
  python - $tag >> $L <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/bench_r2n_{t}.json').read().splitlines() if l.startswith('{')][-1])
    print(f"{t:>16}: ms/step {d['ms_per_step']:.3f}  loss {d['config']['final_loss']:.5f} launches/step {d['gpu_launches']//(d['steps'])}")
except Exception as e:
    print(t, "failed", e)
PY
}
run nofuse   SVDX_GN_BWD_FUSE=0
run fuse     SVDX_GN_BWD_FUSE=1
run nofuse2  SVDX_GN_BWD_FUSE=0
run fuse2    SVDX_GN_BWD_FUSE=1
grep -v "UserWarning\|frombuffer" $L | cut -c1-600 | tail -80
