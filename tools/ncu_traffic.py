#!/usr/bin/env python
"""profiles/ncu_traffic.json: DRAM bytes per launch of the dominant kernel from ncu --set full reports.
usage: python tools/ncu_traffic.py workload=report.ncu-rep ..."""
import csv, io, json, os, subprocess, sys
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
data = json.load(open(out)) if os.path.exists(out) else {}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for arg in sys.argv[1:]:
    wl, rep = arg.split("=")
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def get(k):
        i = hdr.index(k); return float(vals[i].replace(",", "")) * scale[units[i]]
    name = vals[hdr.index("Kernel Name")].split("<")[0].replace("void ", "").split("::")[-1]
    data.setdefault(wl, {})[name] = int(get("dram__bytes_read.sum") + get("dram__bytes_write.sum"))
    data[wl][name + "_source"] = os.path.basename(rep)
json.dump(data, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(data, indent=1))
