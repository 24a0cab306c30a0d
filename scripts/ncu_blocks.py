#!/usr/bin/env python
"""Per-basic-block summary of an ncu report's source page: instruction share, lanes active, stall samples.
  python scripts/ncu_blocks.py gpurun_out/x.ncu-rep [kernel indices...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
which = [int(a) for a in sys.argv[2:]] or None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
ks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; ks.append(cur); continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur is not None and r: cur["rows"].append(r)
for ki, k in enumerate(ks):
    if which is not None and ki not in which: continue
    h = k["hdr"]; iI = h.index("Instructions Executed"); iT = h.index("Thread Instructions Executed"); iS = h.index("# Samples")
    base = int(k["rows"][0][0], 16)
    tot = sum(int(r[iI]) for r in k["rows"]); thr = sum(int(r[iT]) for r in k["rows"]); smp = sum(int(r[iS]) for r in k["rows"])
    print(f"=== kernel {ki} {k['name'][:60]} inst {tot} lanes {thr/max(tot,1):.2f}")
    blk = []
    for r in k["rows"]:
        a = int(r[0], 16) - base; I = int(r[iI]); T = int(r[iT]); S = int(r[iS])
        if blk and abs(blk[-1]["I0"] - I) <= 0.02 * max(I, 1):
            b = blk[-1]; b["n"] += 1; b["I"] += I; b["T"] += T; b["S"] += S; b["end"] = a
        else:
            blk.append(dict(start=a, end=a, n=1, I=I, T=T, S=S, I0=I, first=r[1].strip()[:44]))
    for b in blk:
        if b["I"] / tot > 0.004:
            print(f"{b['start']:05x}-{b['end']:05x} n={b['n']:3d} exec={b['I0']:>10d} share={100*b['I']/tot:5.1f}% lanes={b['T']/max(b['I'],1):5.1f} samples={100*b['S']/max(smp,1):5.1f}%  {b['first']}")
