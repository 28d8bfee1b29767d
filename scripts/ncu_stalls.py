"""Top warp-stall SASS lines of one launch in an .ncu-rep (dev tool): python scripts/ncu_stalls.py rep [skip] [top]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
ci, si, ii = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[2:]):
    try:
        data.append((float(r[si]), n, r))
    except Exception:
        pass
tot = sum(v for v, _, _ in data) or 1.0
agg = collections.Counter()
for v, _, r in data:
    for i in stall:
        try:
            agg[hdr[i]] += float(r[i])
        except Exception:
            pass
print(rows[0][:2], len(data), "lines", int(tot), "samples")
print({k: round(100 * v / tot, 1) for k, v in agg.most_common(8)})
for v, n, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted([(float(r[i] or 0), hdr[i][6:]) for i in stall], reverse=True)[:2]
    print(f"{100 * v / tot:5.1f}% line={n:5d} exec={r[ii]:>8s} {r[ci].strip()[:84]:84s} {st[0][1]}={int(st[0][0])} {st[1][1]}={int(st[1][0])}")
