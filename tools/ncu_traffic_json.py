#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture of tools/profile_step.py: dram__bytes_read.sum +
dram__bytes_write.sum per launch of every voxel kernel (mean over the captured launches), labelled with the git
commit the capture was taken from.  bench.py copies the dominant kernel's figure into `roofline.traffic` only when
the label matches the commit it runs from (or its parent work tree is clean of kernel changes).
usage: ncu_traffic_json.py <rep> <git sha> <out.json>"""
import csv
import json
import subprocess
import sys

rep, sha, out = sys.argv[1:4]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = rows[0]
ki, ri, wi, di = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
units = rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").split("<")[0].replace("cmda::", "")
    b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
    a = acc.setdefault(name, [0.0, 0, 0.0])
    a[0] += b
    a[1] += 1
    a[2] += float(r[di].replace(",", ""))
res = {k: int(v[0] / v[1]) for k, v in acc.items()}
res["_launches"] = {k: v[1] for k, v in acc.items()}
res["_mean_duration_" + units[di]] = {k: v[2] / v[1] for k, v in acc.items()}
res["_git"] = sha
res["_source"] = "ncu --set full --clock-control none (cold caches, serialised), tools/profile_step.py --bins 5 --mode auto --store p4, C2: 16 x 5 M events"
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
