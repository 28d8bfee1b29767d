"""profiles/<tag>_ncu_full.txt from an `ncu --set full` capture of scripts/prof_shapes.py:

    python scripts/summarize_ncu_full.py gpurun_out/prof_r2.ncu-rep gpurun_out/prof_tags.json profiles/r2_ncu_full.txt

One block per profiled kernel: duration, tensor-pipe active %, DRAM bytes and GB/s, L2 -> SM (xbar) bytes, occupancy, registers, and
the achieved algorithmic TFLOP/s or GB/s computed from the tag's FLOPs / bytes (cold-cache single launches under ncu: compare shapes,
not absolute step times)."""
import csv
import json
import re
import subprocess
import sys

PEAK_TF, PEAK_GB = 1355.8, 6581.9


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def main(rep, tags_json, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, pat):
        for h, i in col.items():
            if re.search(pat, h):
                v = num(r[i])
                if v is not None:
                    return v, units[i]
        return None, None

    tags = json.load(open(tags_json))
    kernels = [r for r in data if "cast" not in r[col["Kernel Name"]]]      # the capture is already filtered by -k regex
    k = 0
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, one launch of each hot kernel at its config-2 shape (scripts/prof_shapes.py).\n"
                f"# peaks (MEASURED_PEAKS.json): {PEAK_TF} TFLOP/s sustained bf16, {PEAK_GB} GB/s HBM. Single cold launches: per-shape evidence, not step time.\n")
        for t in tags:
            f.write(f"\n== {t['name']}\n")
            tot_us = 0.0
            for _ in range(t["launches"]):
                if k >= len(kernels):
                    break
                r = kernels[k]
                k += 1
                name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
                dur, du = get(r, r"^gpu__time_duration\.sum$")
                us = dur / 1e3 if du in ("ns", "nsecond") else dur * 1e3 if du in ("ms", "msecond") else dur
                us = max(us, 1e-3)
                tot_us += us
                tens, _ = get(r, r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active")
                if tens is None:
                    tens, _ = get(r, r"sm__pipe_tensor.*cycles_active.*pct")
                rd, ru = get(r, r"^dram__bytes_read\.sum$")
                wr, wu = get(r, r"^dram__bytes_write\.sum$")
                mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                dram = (rd or 0) * mul.get(ru, 1) + (wr or 0) * mul.get(wu, 1)
                xb, xu = get(r, r"^l1tex__m_xbar2l1tex_read_bytes\.sum$")
                xbar = (xb or 0) * mul.get(xu, 1)
                occ, _ = get(r, r"sm__warps_active\.avg\.pct_of_peak_sustained_active")
                regs, _ = get(r, r"^launch__registers_per_thread$")
                grid = r[col["Grid Size"]] if "Grid Size" in col else "?"
                f.write(f"   {name}  grid {grid}  {us:8.1f} us | tensor pipe active {tens if tens is not None else float('nan'):5.1f} % | DRAM {dram / 1e6:8.1f} MB "
                        f"({dram / us / 1e3:6.0f} GB/s) | L2->SM {xbar / 1e6:8.1f} MB ({xbar / us / 1e3:6.0f} GB/s) | warps active {occ if occ is not None else float('nan'):4.1f} % | regs {int(regs) if regs else '?'}\n")
            if t["flops"]:
                tf = t["flops"] / tot_us / 1e6
                f.write(f"   -> algorithmic {t['flops'] / 1e9:.1f} GFLOP in {tot_us:.1f} us = {tf:.0f} TFLOP/s = {tf / PEAK_TF:.3f} of the sustained bf16 peak\n")
            if t["bytes"]:
                gb = t["bytes"] / tot_us / 1e3
                f.write(f"   -> algorithmic {t['bytes'] / 1e6:.1f} MB in {tot_us:.1f} us = {gb:.0f} GB/s = {gb / PEAK_GB:.3f} of the measured HBM peak\n")
    print("wrote", out, "kernels", k, "of", len(kernels))


if __name__ == "__main__":
    main(*sys.argv[1:4])
