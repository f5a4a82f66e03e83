#!/usr/bin/env python3
"""Summarise `nvcc -Xptxas -v` logs: registers / stack / spills per kernel."""
import re, subprocess, sys
for path in sys.argv[1:]:
    log = open(path).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers([^\n]*)", log):
        name, stack, ss, sl, regs, rest = m.groups()
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("amhd::", "").split("(")[0]
        sm = re.search(r"(\d+) bytes smem", rest)
        print(f"{int(regs):4d} regs  stack {int(stack):5d}  spill {ss}/{sl}  {dem[:110]}")
