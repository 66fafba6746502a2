"""Copy the evidence of tools/gpu_final_profile.sh from gpurun_out/final into profiles/<round>/ and refresh profiles/traffic.json."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
src, dst = os.path.join(REPO, "gpurun_out", "final"), os.path.join(REPO, "profiles", rnd)
os.makedirs(dst, exist_ok=True)
def ncu(rep, page, *extra):
    return subprocess.run(["ncu", "-i", rep, "--page", page, *extra], capture_output=True, text=True).stdout


def archive(rep_name, tag, kernel_regex):
    """raw metrics (every captured launch), per-opcode table and details of the kernel matching `kernel_regex`"""
    rep = os.path.join(src, rep_name)
    if not os.path.exists(rep):
        return None
    open(os.path.join(dst, f"final_{tag}_raw.csv"), "w").write(ncu(rep, "raw", "--csv"))
    open("/tmp/src_final.csv", "w").write(ncu(rep, "source", "--csv", "--kernel-name", "regex:" + kernel_regex))
    ops = subprocess.run([sys.executable, os.path.join(REPO, "tools", "ncu_ops.py"), "/tmp/src_final.csv", "1048576"],
                         capture_output=True, text=True).stdout
    open(os.path.join(dst, f"final_{tag}_per_opcode.txt"), "w").write(ops)
    det = ncu(rep, "details")
    open(os.path.join(dst, f"final_{tag}_details.txt"), "w").write("\n".join(l for l in det.splitlines() if l.strip())[:40000])
    return list(csv.reader(open(os.path.join(dst, f"final_{tag}_raw.csv"))))


rows_coop = archive("prof_step_coop.ncu-rep", "step_coop", "step_coop")
rows = archive("prof_step_table.ncu-rep", "step_table", "step_table")
shutil.copy(os.path.join(src, "launches.csv"), os.path.join(dst, "final_launches.csv"))
for f in os.listdir(src):
    if f.startswith("bench_") and f.endswith(".json"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, "final_" + f))
hdr, units = rows[0], rows[1]
r = next(x for x in rows[2:] if "step_table" in x[hdr.index("Kernel Name")])  # the dominant kernel of the default run
g = lambda k: float(r[hdr.index(k)])  # noqa: E731
mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
rd, wr = g("dram__bytes_read.sum") * mult, g("dram__bytes_write.sum") * mult
json.dump({"workload": "cfg3_1m", "envs": 1048576, "kernel": "step_table_kernel", "dram_bytes_per_launch": rd + wr,
           "dram_bytes_read": rd, "dram_bytes_written": wr,
           "source": f"profiles/{rnd}/final_step_table_raw.csv (ncu --set full, one launch)"},
          open(os.path.join(REPO, "profiles", "traffic.json"), "w"), indent=1)
for rws in (rows, rows_coop):
    if not rws:
        continue
    h, u = rws[0], rws[1]
    for r in rws[2:]:
        print("==", r[h.index("Kernel Name")][:40])
        for k in ("gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                  "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                  "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                  "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
                  "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in h:
                print(k, r[h.index(k)], u[h.index(k)])
print("dram read/write MB", rd / 1e6, wr / 1e6)
lrows = list(csv.reader(open(os.path.join(dst, "final_launches.csv"))))
hi = next(i for i, x in enumerate(lrows) if x and x[0] == "ID")
kn, mv = lrows[hi].index("Kernel Name"), lrows[hi].index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for x in lrows[hi + 1:]:
    if len(x) > mv:
        agg[x[kn].split("(")[0][-50:]][0] += 1
        agg[x[kn].split("(")[0][-50:]][1] += float(x[mv])
summary = ["ncu launch list of `bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline` (cold caches, serialised):"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:6]:
    summary.append(f"{v[0]:4d} x {k:55s} {v[1] / v[0] / 1e3:8.1f} us avg {v[1] / 1e3:10.1f} us total")
open(os.path.join(dst, "final_launches_summary.txt"), "w").write("\n".join(summary) + "\n")
print("\n".join(summary))
for f in sorted(os.listdir(dst)):
    if f.startswith("final_bench_") and f.endswith(".json"):
        d = json.loads(open(os.path.join(dst, f)).read().strip().splitlines()[-1])
        print(f, "%.3e" % d["value"], d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"),
              (d.get("roofline_canonical_466B") or {}).get("frac"), (d.get("e2e") or {}).get("value"),
              (d.get("rollout") or {}).get("value"), (d.get("policy_in_loop") or {}).get("value"))
