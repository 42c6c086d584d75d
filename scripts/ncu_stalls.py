#!/usr/bin/env python
"""Stall-reason and opcode mix of one kernel in an ncu report: ncu_stalls.py <rep> <kernel regex>"""
import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1:3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                      ], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
import re
start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and re.search(kern, r[1]))
hdr = rows[start + 1]; ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); opcount = collections.Counter(); opsamp = collections.Counter()
for r in rows[start + 2:]:
    if not r or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name": break
        continue
    toks = r[ci["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    opcount[op] += int(r[ci["Instructions Executed"]] or 0)
    opsamp[op] += int(r[ci["# Samples"]] or 0)
    for s in stalls: tot[s] += int(r[ci[s]] or 0)
T = sum(tot.values()); I = sum(opcount.values()); S = sum(opsamp.values())
print(rows[start][1][:90], "instr", I)
print("stalls %:", [(k.replace('stall_', ''), round(v / T * 100, 1)) for k, v in tot.most_common(10)])
print("ops (instr %, samples %):", [(k, round(v / I * 100, 1), round(opsamp[k] / S * 100, 1)) for k, v in opcount.most_common(20)])
