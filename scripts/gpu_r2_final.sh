#!/bin/bash
# round-2 final evidence on one B200: tests, smoke, the default bench line, configs 4 / 5, kbench, ncu launch lists (cold + warm), ncu --set full
mkdir -p gpurun_out
L=gpurun_out/final.log
: > $L
echo "=== pytest -m gpu" >> $L
timeout 1800 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | grep -v "^$" | cut -c1-300 | tail -6 >> $L
echo "=== smoke" >> $L
timeout 600 python __graft_entry__.py smoke 2>&1 | grep -v Warning | tail -4 >> $L
echo "=== bench default (config 2)" >> $L
timeout 1500 python bench.py > gpurun_out/bench_final_c2.json 2>> gpurun_out/final_err.log
for c in 4 5; do
  echo "=== bench config $c" >> $L
  timeout 1500 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline --no-script-path > gpurun_out/bench_final_c$c.json 2>> gpurun_out/final_err.log
done
python - >> $L <<'PY'
import json
for c in (2, 4, 5):
    try:
        d = json.loads([l for l in open(f'gpurun_out/bench_final_c{c}.json').read().splitlines() if l.startswith('{')][-1])
        print(f"config {c}: ms/step {d['ms_per_step']:.3f} value {d['value']:.1f} roofline {d['roofline']['frac']:.3f} e2e {d['e2e']['value']:.1f} launches {d['gpu_launches']} clocks {d['clocks']}")
        for k in ('script_path', 'vae_encode', 'gpu_eager_baseline', 'cpu_baseline'):
            if k in d: print("   ", k, json.dumps(d[k])[:300])
        for k, v in d.get('roofline_by_family', {}).items(): print("   ", k, round(v['ms_per_step'], 3), "ms  frac", round(v['frac'], 3))
    except Exception as e:
        print("config", c, "failed", e)
PY
echo "=== kbench" >> $L
timeout 900 python scripts/kbench.py gn ln attn wgrad gemm res gnb > gpurun_out/kbench_final.txt 2>&1
tail -3 gpurun_out/kbench_final.txt >> $L
echo "=== ncu launch lists" >> $L
SVDX_SHAPE_LOG=gpurun_out/shapes_final.json timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_final.csv python bench.py --profile-one --warmup 1 --no-graph > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_finalwarm.csv python bench.py --profile-one --warmup 1 --no-graph >> gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tapgemm -c 4000 --csv --log-file gpurun_out/tapgemm_dram_final.csv python bench.py --profile-one --warmup 1 --no-graph >> gpurun_out/ncu_list.log 2>&1
ls -la gpurun_out/launches_final*.csv gpurun_out/tapgemm_dram_final.csv >> $L 2>&1
echo "=== ncu --set full" >> $L
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'tapgemm|attn_|gn_|ln_|adamw|geglu|gemv' -f -o gpurun_out/prof_final python scripts/prof_shapes.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof_final.ncu-rep >> $L 2>&1
python scripts/summarize_ncu_full.py gpurun_out/prof_final.ncu-rep gpurun_out/prof_tags.json gpurun_out/ncu_full_final.txt >> $L 2>&1
rm -f gpurun_out/prof_final.ncu-rep
grep -v "UserWarning\|frombuffer" $L | cut -c1-400 | tail -70
