#!/usr/bin/env python
"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the kernels of ONE force evaluation (between two
launches of the dominant kernel) with their durations and shares.  python tools/launch_list.py <csv> [dominant-kernel-regex]"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    dom = re.compile(sys.argv[2] if len(sys.argv) > 2 else "k_fused_tc")
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        seq.append((r[ki], v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)))
    idx = [i for i, (k, _) in enumerate(seq) if dom.search(k)]
    print("%d launches captured, %d of the dominant kernel" % (len(seq), len(idx)))
    if len(idx) < 3:
        return
    a, b = idx[-3] + 1, idx[-2] + 1          # one full step: everything after a dominant launch up to and including the next
    step = seq[a:b]
    tot = sum(ms for _, ms in step)
    print("| # | kernel | ms | share |\n|---|---|---:|---:|")
    for i, (k, ms) in enumerate(step):
        print("| %d | `%s` | %.4f | %.1f %% |" % (i, re.sub(r"\(.*", "", k)[:80], ms, 100 * ms / tot))
    print("| | **sum (cold-cache, serialised)** | %.3f | |" % tot)


if __name__ == "__main__":
    main()
