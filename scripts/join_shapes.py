"""Join the ncu launch list of one profiled step with the tapgemm shape log of the same step.

    SVDX_SHAPE_LOG=gpurun_out/shapes.json ncu --metrics gpu__time_duration.sum ... python bench.py --profile-one --no-graph
    python scripts/join_shapes.py gpurun_out/launches_X.csv gpurun_out/shapes.json profiles/rN_tapgemm_by_shape.txt

The last len(shape log) tapgemm kernels of the launch list are the profiled step's launches, in order. Output: per-shape
kernel time (cold-cache, serialised by ncu), TFLOP/s, and totals per epilogue kind."""
import collections
import csv
import json
import sys


def us(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v


def main(launch_csv, shape_json, out_path):
    with open(launch_csv) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum" and "tapgemm" in r["Kernel Name"]]
    shapes = json.load(open(shape_json))
    rows = rows[-len(shapes):]
    assert len(rows) == len(shapes), (len(rows), len(shapes))
    agg = collections.OrderedDict()
    for r, s in zip(rows, shapes):
        kern = "pair" if "tapgemm2" in r["Kernel Name"] else "1cta"
        key = (s["M"], s["N"], s["K"], s["taps"], s["conv2d"], s["geglu"], s["a_mn"], s["split_k"], s["res"], s["scales"], s["pre"], s["f32out"], kern,
               s.get("block_n"))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us(r)
    tot = sum(v[1] for v in agg.values())
    with open(out_path, "w") as f:
        f.write(f"# per-shape svdx_tapgemm kernel time of ONE train step (ncu gpu__time_duration, cold cache, serialised); {len(shapes)} launches, "
                f"{tot / 1e3:.2f} ms\n")
        f.write(f"# {'ms':>7s} {'n':>4s} {'us/launch':>9s} {'TFLOP/s':>8s}  shape\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            M, N, K, taps, conv, geglu, amn, sk, res, sc, pre, f32, kern, bn = k
            fl = 2.0 * M * N * K * taps * n
            f.write(f"  {t / 1e3:7.3f} {n:4d} {t / n:9.1f} {fl / t / 1e6:8.1f}  M={M:6d} N={N:5d} K={K:6d} taps={taps} conv2d={conv} geglu={geglu} wgrad={amn} "
                    f"split_k={sk} res={res} scales={sc} pre={pre} f32out={f32} kernel={kern} block_n={bn}\n")
    print("wrote", out_path, "total ms", round(tot / 1e3, 2))


if __name__ == "__main__":
    main(*sys.argv[1:4])
