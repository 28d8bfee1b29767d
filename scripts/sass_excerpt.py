"""profiles/<tag>_sass_excerpt.txt: Blackwell-specific instruction counts per kernel from `cuobjdump -sass` of the shipped library.

    python scripts/sass_excerpt.py profiles/r2_sass_excerpt.txt"""
import collections
import re
import subprocess
import sys

LIB = "svd_xtend_b200/lib/libsvdx_b200.so"
PAT = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
       ("UTMASTG", r"\bUTMASTG"), ("UTMAREDG", r"\bUTMAREDG"), ("UTCBAR", r"\bUTCBAR"), ("HMMA", r"\bHMMA"), ("LDSM", r"\bLDSM"),
       ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("REDG", r"\bREDG|\bRED\.E"), ("MUFU.EX2", r"\bMUFU\.EX2")]


def main(out):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    samples = {}
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PAT:
            if re.search(pat, line):
                per[cur][name] += 1
                if name not in samples and name in ("UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "LDGSTS"):
                    samples[name] = (cur, re.sub(r"\s+", " ", re.sub(r"/\*[0-9a-f]+\*/", "", line)).strip())
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass {LIB} (sm_100a), round-2 final tree: Blackwell-specific instruction counts per kernel\n"
                "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG/UTMAREDG = TMA load / store / reduce-add,\n"
                "# UTCBAR = tcgen05.commit, HMMA = warp-level mma.sync (temporal attention), LDSM = ldmatrix, LDGSTS = cp.async (norm-kernel rings),\n"
                "# SYNCS = mbarrier ops, REDG = red.global\n\n")
        f.write("TOTAL  " + "  ".join(f"{k}={tot[k]}" for k, _ in PAT if tot[k]) + "\n\n")
        for fn, c in per.items():
            if any(c[k] for k in ("UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "LDGSTS")):
                f.write(fn[:150] + "\n    " + "  ".join(f"{k}={c[k]}" for k, _ in PAT if c[k]) + "\n")
        f.write("\n# sample lines\n")
        for k, (fn, line) in samples.items():
            f.write(f"{k:14s} {line[:150]}    <- {fn[:60]}\n")
    print("wrote", out, len(per), "functions")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/r2_sass_excerpt.txt")
