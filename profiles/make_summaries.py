#!/usr/bin/env python
"""Turns the raw files of a capture (gpurun_out/) into the tracked summaries under profiles/.
usage: python profiles/make_summaries.py <tag>      e.g. final -> reads gpurun_out/r02_*_final.*"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "final"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def rows_of(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    return [dict(zip(h, r)) for r in rows[hi + 1:] if len(r) >= len(h)]


def us(v, u):
    v = float(v)
    return v / 1000 if u.startswith("n") else (v if u.startswith("u") else v * 1000)


def kname(full):
    m = re.search(r"&(\w+)", full)
    return m.group(1) if m else full.split("(")[0][:40]


# ---- launch list of one step
seq = [(kname(r["Kernel Name"]), us(r["Metric Value"], r["Metric Unit"]), r["Grid Size"]) for r in rows_of(f"{G}/r02_launches_{tag}.csv") if r["Metric Name"] == "gpu__time_duration.sum"]
idx = [i for i, s in enumerate(seq) if s[0] == "kvec_body"]
step = [s for s in seq[idx[-1]:] if not s[0].startswith("kh_peak") and not s[0].startswith("void at")]
agg = collections.OrderedDict()
for nm, t, g in step:
    a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
with open(f"{P}/r02_launches_{tag}_summary.csv", "w") as f:
    f.write(f"# ncu launch list of ONE step: python bench.py --steps 1 --warmup 1 --no-cpu --no-extra (bzi77, n=98, 4141 solves/step, doubling method), round 2 capture '{tag}'\n")
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES with bench.py's roofline.kernels)\n")
    f.write("kernel,launches,total_us,share\n")
    for nm, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{nm},{c},{t:.1f},{t / tot:.4f}\n")
    f.write(f"# total,{sum(a[0] for a in agg.values())},{tot:.1f},1.0\n# sequence of the step (kernel, us, grid):\n")
    for nm, t, g in step:
        f.write(f"# {nm},{t:.1f},{g}\n")
# ---- DRAM traffic of the dominant kernel
per = collections.defaultdict(dict)
for r in rows_of(f"{G}/r02_zgemm_dram_{tag}.csv"):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r["Metric Unit"]]
    per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"]) * mult
    per[r["ID"]]["grid"] = r["Grid Size"]
single = [d for d in per.values() if d["grid"].startswith("(16564") and len(d) == 3]
totb = [d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in single]
json.dump({"workload": "bzi77", "solves_per_step_per_gpu": 4141, "kernel": "zgemm", "method": "doubling", "dram_bytes_per_launch": sum(totb) / len(totb),
           "launches_averaged": len(single), "algorithmic_bytes_per_launch": 4141 * 3 * 98 * 98 * 16,
           "how": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:zgemm on python bench.py --steps 1 --warmup 1 --no-cpu --no-extra (round 2 capture '{tag}'); "
                  "single-product launches (grid 16564 = 4141 matrices x 4 tiles); algorithmic = read A, B + write C (products that square a matrix read one operand only)"},
          open(f"{P}/r02_dominant_traffic.json", "w"), indent=1)
# ---- ncu --set full summaries
for rep, out, what in ((f"r02_zgemm_{tag}.ncu-rep", "r02_zgemm56u3.txt", "zgemm56u3"), (f"r02_zinv_{tag}.ncu-rep", "r02_zinv_dmma.txt", "zinv_dmma")):
    txt = subprocess.run([sys.executable, f"{P}/ncu_summary.py", f"{G}/{rep}"], capture_output=True, text=True).stdout
    open(f"{P}/{out}", "w").write(f"# ncu --set full --clock-control none --import-source on, one launch of {what} in: python bench.py --steps 1 --warmup 1 --no-cpu --no-extra "
                                  f"(bzi77, n=98, 4141 matrices, doubling method), round 2 capture '{tag}'; summary by profiles/ncu_summary.py\n" + txt)
if os.path.exists(f"{G}/r02_fld_grid_{tag}.ncu-rep"):
    txt = subprocess.run([sys.executable, f"{P}/ncu_summary.py", f"{G}/r02_fld_grid_{tag}.ncu-rep"], capture_output=True, text=True).stdout
    open(f"{P}/r02_fld_grid.txt", "w").write(f"# ncu --set full --clock-control none --import-source on, the fld_grid launch of: python profiles/fields_bench.py 17 "
                                              f"(C5: 9x9 harmonics, 256x256x128 grid, 17 volumes = 13.7 GB written), round 2 capture '{tag}'; summary by profiles/ncu_summary.py\n" + txt)
# ---- bench lines, test logs
for f in (f"r02_bench_{tag}_bzi77.json", f"r02_bench_{tag}_reference_arm.json", f"r02_bench_{tag}_bzi77_full.json", f"r02_pytest_gpu_{tag}.log", f"r02_smoke_{tag}.log", f"r02_accuracy_{tag}.jsonl", f"r02_fields_{tag}_51.jsonl",
          f"r02_fields_{tag}_17.jsonl", f"r02_bench_{tag}_woodpile.json", f"r02_bench_{tag}_suh03.json"):
    if os.path.exists(f"{G}/{f}"):
        open(f"{P}/{f}", "w").write(open(f"{G}/{f}").read())
print("ok", len(step), "launches,", f"{tot / 1000:.2f} ms serialised")
