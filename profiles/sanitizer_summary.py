"""Condense the compute-sanitizer logs of profiles/run_sanitize.sh: tool summaries plus, for racecheck, the distinct
(read site, write site) pairs behind the reported hazards.   python profiles/sanitizer_summary.py <dir with sanitizer_*.log>"""
import collections
import glob
import os
import re
import sys

d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
for path in sorted(glob.glob(os.path.join(d, "sanitizer_*.log")) + glob.glob(os.path.join(d, "r02_memcheck.log")) + glob.glob(os.path.join(d, "r02_racecheck_*.log"))):
    txt = open(path, errors="replace").read()
    print("==", os.path.basename(path))
    for line in txt.splitlines():
        if "SUMMARY" in line or line.startswith(("rollout:", "mcts:", "selfplay:")) or " exit " in line:
            print("   ", line.strip("= ").strip())
    kinds = collections.Counter(re.findall(r"Potential (\w+) hazard", txt))
    if kinds:
        print("    hazards displayed by kind:", dict(kinds))
        sites = collections.Counter()
        rec = re.findall(r"(Read|Write) Thread \([^)]*\) at ([^\n]+?)\+0x[0-9a-f]+ in ([\w./]+:\d+)", txt)
        for i in range(0, len(rec) - 1, 2):
            sites[(rec[i][0] + " " + rec[i][1] + " " + rec[i][2], rec[i + 1][0] + " " + rec[i + 1][1] + " " + rec[i + 1][2])] += 1
        for (a, b), c in sites.most_common():
            print(f"    {c:6d} x  {a}  <->  {b}")
