"""Aggregate an `ncu --page source --csv` dump per opcode: executions, L1 wavefronts, L2 sectors, stall samples.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_ops.py src.csv <envs per launch> [launches]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
envs = int(sys.argv[2]) * (int(sys.argv[3]) if len(sys.argv) > 3 else 1)
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
I = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0] * 7)
tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    toks = r[I["Source"]].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    parts = op.split(".")
    key = parts[0]
    if key in ("LDG", "STG", "LDS", "STS"):
        key += "." + next((p for p in parts if p in ("64", "128", "U8", "S8", "U16")), "32")

    def f(k):
        try:
            return int(r[I[k]])
        except Exception:
            return 0

    a = agg[key]
    for j, k in enumerate(("Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                           "L1 Tag Requests Global", "L2 Theoretical Sectors Global",
                           "L2 Theoretical Sectors Global Ideal", "# Samples")):
        a[j] += f(k)
    tot += f("# Samples")
print(f"{'op':10s} {'exec/env':>9s} {'smem_wf':>8s} {'ideal':>8s} {'l1tag':>8s} {'l2sect':>8s} {'ideal':>8s} {'stall%':>7s}")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][6])[:30]:
    print(f"{op:10s} {a[0]/envs:9.3f} {a[1]/envs:8.3f} {a[2]/envs:8.3f} {a[3]/envs:8.3f} {a[4]/envs:8.3f} {a[5]/envs:8.3f} {100*a[6]/max(tot,1):7.1f}")
s = [sum(a[j] for a in agg.values()) / envs for j in range(7)]
print(f"{'TOTAL':10s} {s[0]:9.3f} {s[1]:8.3f} {s[2]:8.3f} {s[3]:8.3f} {s[4]:8.3f} {s[5]:8.3f}")
