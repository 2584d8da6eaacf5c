#!/usr/bin/env python
"""Writes profiles/ncu_traffic.json (read by bench.py for roofline.traffic) from an `ncu --set full` report of the
dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum per launch, with the kernel name, the number of problems
the captured launch solved and the git SHA of the tree the capture was taken from.
usage: ncu_traffic.py report.ncu-rep CONFIG problems_per_launch [git_sha]"""
import csv, json, os, subprocess, sys
rep, cfg, nprob = sys.argv[1], sys.argv[2], int(sys.argv[3])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sha = sys.argv[4] if len(sys.argv) > 4 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = sum(float(d[m]) * scale[u[m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
cur = json.load(open(path)) if os.path.exists(path) else {}
cur[cfg] = {"dram_bytes": int(tot), "problems_per_launch": nprob, "kernel": d["Kernel Name"].split("(")[0][:80],
            "git_sha": sha, "report": os.path.basename(rep), "duration_ms": float(d["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "s": 1e3}.get(u["gpu__time_duration.sum"], 1)}
json.dump(cur, open(path, "w"), indent=1)
print(json.dumps(cur[cfg]))
