#!/usr/bin/env python
"""profiles/<name>.md from an ncu report: headline metrics + the hottest source lines (instructions, active lanes, stall samples).
usage: python tools/make_profile_md.py report.ncu-rep profiles/name.md "title / command" """
import subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
summ = subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout
lines = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "40"], capture_output=True, text=True).stdout
with open(out, "w") as f:
    f.write(f"# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` (one launch, after 3 warm-up launches), read with\n"
            f"`tools/ncu_summary.py` / `tools/ncu_lines.py`.  Times under ncu are not bench values.\n\n## Headline metrics\n\n```\n{summ}```\n\n"
            f"## Hottest source lines (share of warp instructions, share of stall samples, average active lanes)\n\n```\n{lines}```\n")
print("wrote", out)
