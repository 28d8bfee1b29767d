#!/bin/bash
# GroupNorm-backward sums fused into the dgrad epilogue: parity, then in-step A/B on the same box
mkdir -p gpurun_out
L=gpurun_out/r2k.log
: > $L
echo "=== pytest gnb" >> $L
timeout 900 python -m pytest tests/test_tapgemm_gpu.py tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "backward_sums or groupnorm" 2>&1 | grep -v "^$" | cut -c1-600 | tail -30 >> $L
echo "=== full gpu tests" >> $L
timeout 1800 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | cut -c1-400 | tail -25 >> $L
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-script-path --no-families"
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 $B > gpurun_out/bench_r2k_$tag.json 2>> gpurun_out/r2k_err.log
  python - $tag >> $L <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/bench_r2k_{t}.json').read().splitlines() if l.startswith('{')][-1])
    print(f"{t:>16}: ms/step {d['ms_per_step']:.3f}  loss {d['config']['final_loss']:.5f} launches/step {d['gpu_launches']//(d['steps'])}")
except Exception as e:
    print(t, "failed", e)
PY
}
run nofuse   SVDX_GN_BWD_FUSE=0
run fuse     SVDX_GN_BWD_FUSE=1
run nofuse2  SVDX_GN_BWD_FUSE=0
run fuse2    SVDX_GN_BWD_FUSE=1
run fuse_nowide SVDX_GN_BWD_FUSE=1 SVDX_WIDE=0
run nofuse_nowide SVDX_GN_BWD_FUSE=0 SVDX_WIDE=0
grep -v "UserWarning\|frombuffer" $L | cut -c1-600 | tail -80
