#!/usr/bin/env python
"""Per-kernel stall / opcode breakdown from the SASS source page of an .ncu-rep.
usage: python tools/ncu_stalls.py rep.ncu-rep [kernel-substring] [ntop]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None and row:
        cur["rows"].append(row)
for b in blocks:
    if want not in b["name"]:
        continue
    hdr, data = b["rows"][0], b["rows"][1:]
    ix = {h: i for i, h in enumerate(hdr)}
    S, src, ie = ix["# Samples"], ix["Source"], ix["Instructions Executed"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[S] or 0) for r in data)
    print("==", b["name"][:90])
    print("static instr", len(data), "samples", tot, "warp instr executed", sum(int(r[ie] or 0) for r in data))
    agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
    print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
    byop = collections.Counter()
    for r in data:
        p = r[src].split()
        if not p:
            continue
        op = p[1] if p[0].startswith("@") else p[0]
        byop[op.split(".")[0]] += int(r[ie] or 0)
    print("executed by opcode:", byop.most_common(16))
    for r in sorted(data, key=lambda r: -int(r[S] or 0))[:ntop]:
        print("  %5s %8s  %-58s %s" % (r[S], r[ie], r[src][:58], {h[6:]: r[ix[h]] for h in stalls if int(r[ix[h]] or 0) > 0.25 * int(r[S] or 1)}))
