#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines:
   python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> <object.o> <mangled-name> [top=25]
The report's SASS listing (program order) is aligned with `nvdisasm -g` of the same kernel (line-info
annotations) by instruction index; samples are summed per (file, line) including inlined callers' lines."""
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

import pandas as pd


def main():
    rep, kern, obj, mangled = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout.split("\n")
    hdr = [i for i, l in enumerate(raw) if l.startswith('"Address"')]
    df = pd.read_csv(io.StringIO("\n".join(raw[hdr[0]:(hdr[1] - 1 if len(hdr) > 1 else len(raw))])))
    n = pd.to_numeric(df["# Samples"], errors="coerce").fillna(0).to_numpy()
    ops = [s.strip() for s in df["Source"]]
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(sass) if l.startswith(".text." + mangled + ":"))
    lines = []       # per instruction: (file, line)
    cur = ("?", 0)
    for l in sass[start + 1:]:
        if l.startswith(".text.") or l.startswith("//-----"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    print("instructions: ncu %d, nvdisasm %d" % (len(ops), len(lines)))
    k = min(len(ops), len(lines))
    agg = defaultdict(float)
    for i in range(k):
        agg[lines[i]] += n[i]
    tot = n.sum()
    src = {}
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        if f not in src:
            p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", f)
            src[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
        print("%5.2f%%  %s:%d  %s" % (100 * v / tot, f, ln, text))


if __name__ == "__main__":
    main()
