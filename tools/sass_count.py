#!/usr/bin/env python
"""Static SASS statistics of an object file / library: instructions per kernel and opcode histogram.
Usage: python tools/sass_count.py rl-itg_b200/csrc/_obj/torm.o [kernel-substring]"""
import collections
import re
import subprocess
import sys


def main():
    out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
    want = sys.argv[2] if len(sys.argv) > 2 else None
    name, stats = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:110]
            stats[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            stats[name][m.group(1).split(".")[0]] += 1
    for k, c in sorted(stats.items(), key=lambda kv: sum(kv[1].values())):
        if want and want not in k:
            continue
        print("%6d  %s" % (sum(c.values()), k))
        if want:
            print("        ", c.most_common(24))


if __name__ == "__main__":
    main()
