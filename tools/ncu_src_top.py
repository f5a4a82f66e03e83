#!/usr/bin/env python3
"""Top stall-sample instructions of an `ncu --page source --csv` export (one kernel): index, SASS, samples, executions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))[2:]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(int(r[2]) for r in rows)
print("total samples", tot, "instructions", len(rows))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i][2]))[:top]
marks = ('BAR', 'MEMBAR', 'SYNCS', 'CCTL', 'ERRBAR')
for i, r in enumerate(rows):
    if i in idx or (any(m in r[1] for m in marks) and int(r[2]) > 0):
        print(i, r[1].strip()[:72], r[2], f"{100 * int(r[2]) / tot:.1f}%", "exec", r[5], "thr", r[8])
