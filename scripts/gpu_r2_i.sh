#!/bin/bash
# (historical A/B: the register-array GroupNorm / LayerNorm kernels selected by SVDX_GN_RING=0 / SVDX_LN_RING=0 / SVDX_LIB alt builds were removed after this comparison)
# in-step A/B of the ring norm kernels (same box): whole-step time per variant, warm-L2 kbench, ncu launch list without cache flushes
mkdir -p gpurun_out
L=gpurun_out/r2i.log
: > $L
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-script-path --no-families"
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 $B > gpurun_out/bench_r2i_$tag.json 2>> gpurun_out/r2i_err.log
  python - $tag >> $L <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/bench_r2i_{t}.json').read().splitlines() if l.startswith('{')][-1])
    print(f"{t:>16}: ms/step {d['ms_per_step']:.3f}  loss {d['config']['final_loss']:.5f}")
except Exception as e:
    print(t, "failed", e)
PY
}
run base     SVDX_GN_RING=0 SVDX_LN_RING=0
run gnring   SVDX_GN_RING=1 SVDX_LN_RING=0
run lnring   SVDX_GN_RING=0 SVDX_LN_RING=1
run both     SVDX_GN_RING=1 SVDX_LN_RING=1
run both_cps2 SVDX_GN_RING_CPS=2
run both_cps3 SVDX_GN_RING_CPS=3
run both_cps6 SVDX_GN_RING_CPS=6
run both_ln3 SVDX_GN_RING_CPS=3 SVDX_LN_RING_CPS=3
run base2    SVDX_GN_RING=0 SVDX_LN_RING=0
echo "--- warm kbench, base" >> $L
KBENCH_WARM=1 SVDX_GN_RING=0 SVDX_LN_RING=0 timeout 200 python scripts/kbench.py gn ln 2>&1 | grep "M=" >> $L
echo "--- warm kbench, ring cps3" >> $L
KBENCH_WARM=1 SVDX_GN_RING_CPS=3 timeout 200 python scripts/kbench.py gn ln 2>&1 | grep "M=" >> $L
for v in base ring; do
  if [ $v == base ]; then E="SVDX_GN_RING=0 SVDX_LN_RING=0"; else E="SVDX_GN_RING_CPS=3"; fi
  env $E timeout 900 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/launches_warm_$v.csv python bench.py --profile-one --no-graph > gpurun_out/ncu_warm_$v.log 2>&1
done
grep -v "UserWarning\|frombuffer" $L | cut -c1-250 | tail -150
