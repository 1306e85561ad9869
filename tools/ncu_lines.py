#!/usr/bin/env python
"""Per-source-line totals (warp instructions executed, stall samples) of one kernel in an .ncu-rep captured with
--import-source on.  usage: python tools/ncu_lines.py rep kernel-regex [top]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kre, "--launch-count", "1"], capture_output=True, text=True).stdout
    inst, samp, text = collections.Counter(), collections.Counter(), {}
    cur_file, hdr = None, None
    for row in csv.reader(io.StringIO(raw)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1].split("/")[-1]
            continue
        if row[0] == "Line No":
            hdr = row
            continue
        if hdr is None or cur_file is None or not row[0].isdigit():
            continue
        d = dict(zip(hdr, row))
        key = (cur_file, int(row[0]))
        try:
            inst[key] += int(d.get("Instructions Executed") or 0)
            samp[key] += int(d.get("# Samples") or 0)
        except ValueError:
            pass
        text[key] = row[1].strip()[:110]
    tot_i, tot_s = sum(inst.values()), sum(samp.values())
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for key, n in inst.most_common(top):
        print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100.0 * n / max(tot_i, 1), 100.0 * samp[key] / max(tot_s, 1), key[0], key[1], text[key]))


if __name__ == "__main__":
    main()
