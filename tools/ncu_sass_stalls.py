#!/usr/bin/env python
"""Per-SASS-instruction stall samples of one kernel launch in an .ncu-rep (captured with --set full --import-source on):
totals by stall reason, then the most-sampled instructions with their position in the instruction stream and the
reasons.  This is the view that showed the serialised load round trips at the head of the PIRL env kernel (round 2).
usage: python tools/ncu_sass_stalls.py rep kernel-regex [top] [launch-skip]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    hdr, rows = None, []
    for r in csv.reader(io.StringIO(raw)):
        if len(r) > 5 and r[0] == "Address":
            if hdr is not None:
                break
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 5:
            continue
        rows.append(dict(zip(hdr, r)))
    reasons = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    tot, n = collections.Counter(), 0
    per = []
    for i, d in enumerate(rows):
        try:
            s = int(d["# Samples"])
        except ValueError:
            continue
        n += s
        rs = {}
        for k in reasons:
            try:
                v = int(d[k])
            except ValueError:
                v = 0
            tot[k] += v
            if v:
                rs[k[6:]] = v
        per.append((s, i, d["Source"].strip()[:78], rs))
    print("%d instructions, %d samples" % (len(rows), n))
    for k, v in tot.most_common(10):
        print("  %-26s %7d %5.1f%%" % (k, v, 100.0 * v / max(n, 1)))
    for s, i, src, rs in sorted(per, reverse=True)[:top]:
        main_r = ", ".join("%s %d" % kv for kv in sorted(rs.items(), key=lambda kv: -kv[1])[:2])
        print("%5.1f%%  #%-5d %-78s %s" % (100.0 * s / max(n, 1), i, src, main_r))


if __name__ == "__main__":
    main()
