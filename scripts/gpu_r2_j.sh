#!/bin/bash
# 256x320 CTA-pair tiles: parity, tile-width sweep, in-step A/B (same box) together with the norm-ring CTA counts
mkdir -p gpurun_out
L=gpurun_out/r2j.log
: > $L
echo "=== pytest wide tiles" >> $L
timeout 900 python -m pytest tests/test_tapgemm_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "wide" 2>&1 | grep -v "^$" | cut -c1-400 | tail -25 >> $L
echo "=== kbench gemm" >> $L
timeout 900 python scripts/kbench.py gemm 2>&1 | grep -v Warning | tail -40 >> $L
echo "=== full gpu tests" >> $L
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x 2>&1 | grep -v "^$" | cut -c1-300 | tail -8 >> $L
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-script-path --no-families"
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 $B > gpurun_out/bench_r2j_$tag.json 2>> gpurun_out/r2j_err.log
  python - $tag >> $L <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/bench_r2j_{t}.json').read().splitlines() if l.startswith('{')][-1])
    print(f"{t:>16}: ms/step {d['ms_per_step']:.3f}  loss {d['config']['final_loss']:.5f}")
except Exception as e:
    print(t, "failed", e)
PY
}
run nowide_g2l3  SVDX_WIDE=0 SVDX_GN_RING_CPS=2 SVDX_LN_RING_CPS=3
run wide_g2l3    SVDX_WIDE=1 SVDX_GN_RING_CPS=2 SVDX_LN_RING_CPS=3
run wide1280_g2l3 SVDX_WIDE=1 SVDX_WIDE_MIN_K=1280 SVDX_GN_RING_CPS=2 SVDX_LN_RING_CPS=3
run wide_g2l2    SVDX_WIDE=1 SVDX_GN_RING_CPS=2 SVDX_LN_RING_CPS=2
run wide_g2l4    SVDX_WIDE=1 SVDX_GN_RING_CPS=2 SVDX_LN_RING_CPS=4
run wide_g1l3    SVDX_WIDE=1 SVDX_GN_RING_CPS=1 SVDX_LN_RING_CPS=3
run wide_g3l3    SVDX_WIDE=1 SVDX_GN_RING_CPS=3 SVDX_LN_RING_CPS=3
run nowide_g2l3b SVDX_WIDE=0 SVDX_GN_RING_CPS=2 SVDX_LN_RING_CPS=3
grep -v "UserWarning\|frombuffer" $L | cut -c1-400 | tail -100
