#!/usr/bin/env python
"""Per-source-line instruction counts for one kernel of an ncu report.

    python scripts/ncu_lines.py <report.ncu-rep> <kernel regex> <lib.so> [top N]

Joins `ncu --page source --print-source sass` (per-SASS-instruction executed
counts and stall samples) with `nvdisasm -g` line info of the same cubin
(instruction order is identical), and prints the hottest source lines.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kre, lib = sys.argv[1:4]
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                          ], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # first kernel instance only
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and re.search(kre, r[1]))
    kname = rows[start][1]
    hdr = rows[start + 1]
    body = []
    for r in rows[start + 2:]:
        if r and r[0] == "Kernel Name":
            break
        if r and r[0].startswith("0x"):
            body.append(r)
    ci = {h: i for i, h in enumerate(hdr)}
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    dis = ""      # the library holds one cubin per translation unit
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        dis += "\n" + subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # find the function whose demangled name matches: use the template signature in kname
    funcs = re.split(r"\n//-+ \.text\.", dis)
    pick = None
    want = re.sub(r"[^A-Za-z0-9_]", "", kname.split("(")[0].split("::")[-1].split("<")[0])
    targs = re.findall(r"\((bool|int)\)(\d+)", kname)
    pat = ".*".join(f"L{'b' if t == 'bool' else 'i'}{v}E" for t, v in targs)
    for f in funcs[1:]:
        name = f.split(" ", 1)[0]
        if want in name and (not targs or re.search(pat, name)):
            pick = f
            break
    if pick is None:
        print("function not found in cubin for", kname)
        return
    lines = []      # innermost (file, line) per instruction
    chains = []     # full inline chain per instruction, innermost first
    chain = []
    fresh = True
    for ln in pick.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                chain = []
                fresh = False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(chain[0] if chain else ("?", 0))
            chains.append(list(chain))
            fresh = True
    n = min(len(lines), len(body))
    if len(lines) != len(body):
        print(f"warning: {len(lines)} disasm instrs vs {len(body)} profiled", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for i in range(n):
        r = body[i]
        ie = int(r[ci["Instructions Executed"]] or 0)
        sm = int(r[ci["# Samples"]] or 0)
        ti = int(r[ci["Thread Instructions Executed"]] or 0)
        a = agg[lines[i]]
        a[0] += ie; a[1] += sm; a[2] += ti
        tot_i += ie; tot_s += sm
    print(kname)
    print(f"total warp instr {tot_i:,}  samples {tot_s:,}")
    # ---- by enclosing function of the main .cu (phase view) ----
    cu = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kaolin_b200", "csrc", "dibr_b200.cu")
    fn_at = {}
    if os.path.exists(cu):
        cur_fn = "?"
        for no, text in enumerate(open(cu).read().splitlines(), 1):
            m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|static|inline|int |void |size_t ).*?([A-Za-z_0-9]+)\(", text)
            if m and not text.startswith(" "):
                cur_fn = m.group(1)
            fn_at[no] = cur_fn
    phase = collections.defaultdict(lambda: [0, 0])
    site = collections.defaultdict(lambda: [0, 0])
    for i in range(n):
        r = body[i]
        ie = int(r[ci["Instructions Executed"]] or 0)
        sm_ = int(r[ci["# Samples"]] or 0)
        fr = [c for c in chains[i] if c[0] == "dibr_b200.cu"]
        # outermost frame is the kernel; the one before it (if any) is the phase function
        ph = fn_at.get(fr[-2][1], "?") if len(fr) >= 2 else (fn_at.get(fr[-1][1], "?") if fr else "?")
        phase[ph][0] += ie; phase[ph][1] += sm_
        key = fr[0] if fr else ("?", 0)   # innermost frame inside the .cu = call site of the math
        site[key][0] += ie; site[key][1] += sm_
    print("-- by phase function --")
    for k, (ie, sm_) in sorted(phase.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:28s} instr {ie / max(tot_i,1) * 100:5.1f}%  samples {sm_ / max(tot_s,1) * 100:5.1f}%")
    print("-- by innermost line inside dibr_b200.cu --")
    srcl = open(cu).read().splitlines() if os.path.exists(cu) else []
    for (f, l), (ie, sm_) in sorted(site.items(), key=lambda kv: -kv[1][1])[:topn]:
        text = srcl[l - 1].strip()[:80] if 0 < l <= len(srcl) else ""
        print(f"  {l:<5d} instr {ie / max(tot_i,1) * 100:5.1f}%  samples {sm_ / max(tot_s,1) * 100:5.1f}% | {text}")
    return
    src_cache = {}
    for (f, l), (ie, sm, ti) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
        path = os.path.join(os.path.dirname(os.path.abspath(lib)), f)
        if path not in src_cache and os.path.exists(path):
            src_cache[path] = open(path).read().splitlines()
        text = src_cache.get(path, [""] * (l + 1))[l - 1].strip()[:70] if path in src_cache and l <= len(src_cache[path]) else ""
        print(f"{f}:{l:<5d} instr {ie / max(tot_i, 1) * 100:5.1f}%  samples {sm / max(tot_s, 1) * 100:5.1f}%  thr/inst {ti / max(ie, 1):5.1f} | {text}")


if __name__ == "__main__":
    main()
