"""Turn the ncu artefacts of scripts/profile_round.sh into the small tracked files under profiles/.

  python scripts/summarise_profile.py r05

reads  gpurun_out/<tag>_prof_fp16_P4096.ncu-rep, gpurun_out/<tag>_launches_fp16_P4096.csv
writes profiles/<tag>_prof_fp16_P4096.csv (ncu --page raw), profiles/<tag>_launches_fp16_P4096.csv and prints the
per-kernel DRAM bytes / durations (for profiles/traffic.json and profiles/README.md)."""
import csv
import io
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_fp16_P4096.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
out = os.path.join(ROOT, "profiles", f"{tag}_prof_fp16_P4096.csv")
open(out, "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
for r in rows[2:]:
    print("==", r[col["Kernel Name"]][:70], r[col["Block Size"]], r[col["Grid Size"]])
    for w in want:
        if w in col:
            print(f"   {w:62s} {r[col[w]]} {units[col[w]]}")
lst = os.path.join(ROOT, "gpurun_out", f"{tag}_launches_fp16_P4096.csv")
if os.path.exists(lst):
    shutil.copy(lst, os.path.join(ROOT, "profiles", f"{tag}_launches_fp16_P4096.csv"))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    lines = [l for l in open(lst) if l.startswith('"')]
    rd = csv.reader(lines)
    h = next(rd)
    ci = {x: i for i, x in enumerate(h)}
    for r in rd:
        if len(r) <= ci["Metric Value"]:
            continue
        name = r[ci["Kernel Name"]].split("(")[0].split("<")[0]
        tot[name] += float(r[ci["Metric Value"]].replace(",", ""))
        cnt[name] += 1
    s = sum(tot.values())
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"   {k:40s} {cnt[k]:5d} launches  {100 * tot[k] / s:6.2f} % of kernel time")
