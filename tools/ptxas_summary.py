"""Summarise ptxas -v output in build.log: registers and spills per kernel."""
import re, subprocess, sys, os
log = open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "stochastic-epidemic-modelling_b200", "build.log")).read()
rows = []
for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    rows.append((name, int(m.group(5)), int(m.group(3)), int(m.group(4))))
pat = sys.argv[2] if len(sys.argv) > 2 else ""
for n, r, ss, sl in rows:
    if pat in n:
        print(f"{r:4d} regs  spill st/ld {ss:5d}/{sl:5d}  {n[:110]}")
