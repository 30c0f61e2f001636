#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: warp instructions executed, average active threads, stall samples.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None
agg = collections.OrderedDict()
tot_inst = tot_thr = tot_samp = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] not in ("", ) and r[2] == "-":          # a source-line summary row
        try:
            line = int(r[0]); samp = int(r[6]); inst = int(r[7]); thr = int(r[8])
        except ValueError:
            continue
        key = (cur_file, line, r[1].strip()[:90])
        a = agg.setdefault(key, [0, 0, 0]); a[0] += inst; a[1] += thr; a[2] += samp
        tot_inst += inst; tot_thr += thr; tot_samp += samp
print(f"total warp-inst {tot_inst:,}  avg active threads {tot_thr/max(tot_inst,1):.2f}  samples {tot_samp:,}")
print(f"{'file:line':28} {'inst%':>6} {'samp%':>6} {'thr':>5}  source")
for (f, l, s), (inst, thr, samp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{f+':'+str(l):28} {100*inst/tot_inst:6.2f} {100*samp/max(tot_samp,1):6.2f} {thr/max(inst,1):5.1f}  {s}")
