"""Turn the scratch captures under gpurun_out/ into the committed round summaries under profiles/.

    python scripts/summarize_profiles.py r1
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def us(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v


def launch_list():
    src = os.path.join(G, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    idx = [i for i, r in enumerate(rows) if "adamw_state_kernel" in r["Kernel Name"] or "adamw_kernel" in r["Kernel Name"]]
    seg = rows[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else rows
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in seg:
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", r["Kernel Name"]))
        tot[name] += us(r)
        cnt[name] += 1
    T = sum(tot.values())
    with open(os.path.join(P, f"{tag}_launch_list_step.txt"), "w") as f:
        f.write(f"# one train step (between two adamw launches) of `python bench.py --profile-one --no-graph` under\n"
                f"# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)\n"
                f"# launches {len(seg)}  total {T / 1e3:.2f} ms\n")
        f.write(f"{'ms':>9s} {'share':>7s} {'n':>6s} {'avg_us':>9s}  kernel\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"{v / 1e3:9.3f} {100 * v / T:6.2f}% {cnt[k]:6d} {v / cnt[k]:9.1f}  {k}\n")
    print("wrote launch list:", len(seg), "launches", round(T / 1e3, 2), "ms")


def raw_metrics(rep, out, pick):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        return
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    keep = [i for i, h in enumerate(hdr) if any(re.search(p, h) for p in pick)]
    with open(os.path.join(P, out), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; source {rep}\n")
        for d in rows[2:]:
            f.write("----\n")
            for i in keep:
                if d[i] not in ("", "n/a"):
                    f.write(f"{hdr[i]} [{units[i]}] = {d[i]}\n")
    print("wrote", out)


def source_stalls(rep, out, top=40):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        return
    txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[1]
    ci, si, ii = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        try:
            data.append((float(r[si]), r))
        except Exception:
            pass
    tot = sum(v for v, _ in data) or 1.0
    agg = collections.Counter()
    for v, r in data:
        for i in stall:
            try:
                agg[hdr[i]] += float(r[i])
            except Exception:
                pass
    with open(os.path.join(P, out), "w") as f:
        f.write(f"# source-level warp-stall sampling, {rep}; {len(data)} SASS lines, {int(tot)} samples\n# stall reasons: "
                + ", ".join(f"{k}={100 * v / tot:.1f}%" for k, v in agg.most_common(8)) + "\n")
        for v, r in sorted(data, key=lambda x: -x[0])[:top]:
            f.write(f"{100 * v / tot:5.1f}%  exec={r[ii]:>9s}  {r[ci].strip()}\n")
    print("wrote", out)


def dram_traffic():
    """sum of dram bytes / time over every tapgemm launch of one step (cheap 3-metric ncu pass)"""
    src = os.path.join(G, f"tapgemm_dram_{tag}.csv")
    if not os.path.exists(src):
        return
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(dict)
    order = []
    for r in csv.DictReader(lines):
        k = r["ID"]
        if k not in per:
            order.append(k)
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        name = r["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        else:
            v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        per[k][name] = v
        per[k]["kernel"] = re.sub(r"\(.*", "", r["Kernel Name"])
    ids = order
    # one steady-state step = the last 1/2 of the captured launches is safest; bench --profile-one runs warm-up(3)+1 steps
    n_step = len(ids) // 4 if len(ids) >= 8 else len(ids)
    sel = ids[-n_step:]
    rd = sum(per[k].get("dram__bytes_read.sum", 0) for k in sel)
    wr = sum(per[k].get("dram__bytes_write.sum", 0) for k in sel)
    tm = sum(per[k].get("gpu__time_duration.sum", 0) for k in sel)
    out = {"kernel": "svdx::tapgemm_kernel + svdx::tapgemm2_kernel", "launches_per_step": n_step, "dram_read_bytes_per_step": rd,
           "dram_write_bytes_per_step": wr, "traffic_bytes_per_launch": (rd + wr) / max(n_step, 1), "ncu_time_ms_per_step": tm / 1e3,
           "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tapgemm; tapgemm_dram_{tag}.csv, "
                     f"last {n_step} of {len(ids)} captured launches"}
    json.dump(out, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
    print("wrote traffic:", out)


PICK = [r"^Kernel Name$", r"^Grid Size$", r"^Block Size$", r"gpu__time_duration\.sum$", r"dram__bytes_(read|write)\.sum$",
        r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"sm__pipe_tensor_cycles_active.*pct", r"sm__warps_active\.avg\.pct",
        r"launch__registers_per_thread$", r"l1tex__m_xbar2l1tex_read_bytes\.sum(\.per_second)?$", r"lts__t_bytes\.sum$",
        r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"launch__shared_mem_per_block_dynamic", r"sm__inst_executed_pipe_tensor"]

if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    launch_list()
    dram_traffic()
    raw_metrics(f"prof_tapgemm_{tag}.ncu-rep", f"{tag}_tapgemm_ncu_full.txt", PICK)
    raw_metrics(f"prof_attn_fwd_{tag}.ncu-rep", f"{tag}_attn_fwd_ncu_full.txt", PICK)
    source_stalls(f"prof_attn_fwd_{tag}.ncu-rep", f"{tag}_attn_fwd_source_stalls.txt")
    source_stalls(f"prof_attn_dq_{tag}.ncu-rep", f"{tag}_attn_bwd_dq_source_stalls.txt")
    gt = os.path.join(G, "gemm_table.json")
    if os.path.exists(gt):
        t = json.load(open(gt))
        with open(os.path.join(P, f"{tag}_tapgemm_shapes.txt"), "w") as f:
            f.write("# per-shape tapgemm time inside one eager train step (CUDA events around every launch, bench.py SVDX_GEMM_TABLE)\n")
            f.write(f"# total {sum(r['ms'] for r in t):.2f} ms, {sum(r['n'] for r in t)} launches\n")
            for r in t:
                f.write(f"{r['ms']:7.3f} ms n={r['n']:3d} {r['tflops']:7.1f} TFLOP/s  M={r['M']:6d} N={r['N']:5d} K={r['K']:6d} taps={r['taps']} "
                        f"conv2d={r['mode']} geglu={int(r['geglu'])} wgrad={int(r['wgrad'])} split_k={r['split_k']}\n")
        print("wrote shapes table")
