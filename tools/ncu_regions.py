#!/usr/bin/env python
"""Attribute ncu warp-stall samples to the inline call path of every SASS instruction:
   python tools/ncu_regions.py <report.ncu-rep> <kernel-regex> <object.o> <mangled-kernel-name> [depth=3] [top=40]
`nvdisasm -gi` annotates every instruction with its chain of inlined call sites; each source line is mapped to the
function that contains it (function start lines parsed from the .cuh files), which gives a path such as
k_fused_tc > t_body > tc_tp_backward > vec_load.  Samples (and the stall-reason mix) are summed per path prefix.
Development aid; the persistent fused kernel makes one capture a complete profile of all phases."""
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

import pandas as pd

FUNC_RE = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:__global__|__device__|static|inline|__host__|ALG_NI|__forceinline__|\s)+"
                     r"[\w:<>\*&\s]*?\b(\w+)\s*\(")


def function_table(path):
    """[(start_line, name)] of the function definitions of a source file (good enough for our own headers)"""
    out = []
    if not os.path.exists(path):
        return out
    pending_template = False
    for i, l in enumerate(open(path, errors="ignore").read().split("\n"), 1):
        s = l.strip()
        if s.startswith("template") and "(" not in s:
            pending_template = True
            continue
        m = re.match(r"^(?:template\s*<.*>\s*)?(?:__global__|__device__|__host__)[^;]*?\b(\w+)\s*\(", s)
        if m and not s.endswith(";"):
            out.append((i, m.group(1)))
        pending_template = False
    return out


def main():
    rep, kern, obj, mangled = sys.argv[1:5]
    depth = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout.split("\n")
    hdr = [i for i, l in enumerate(raw) if l.startswith('"Address"')]
    df = pd.read_csv(io.StringIO("\n".join(raw[hdr[0]:(hdr[1] - 1 if len(hdr) > 1 else len(raw))])))
    n = pd.to_numeric(df["# Samples"], errors="coerce").fillna(0).to_numpy()
    stalls = [c for c in df.columns if c.startswith("stall_") and "Not Issued" not in c]
    st = {c: pd.to_numeric(df[c], errors="coerce").fillna(0).to_numpy() for c in stalls}
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(sass) if l.startswith(".text." + mangled + ":"))
    tables = {}

    def fn_of(f, ln):
        if f not in tables:
            tables[f] = function_table(f)
        name = None
        for s0, nm in tables[f]:
            if s0 <= ln:
                name = nm
            else:
                break
        return name or os.path.basename(f)

    paths = []
    chain = []
    cur = ("?",)
    for l in sass[start + 1:]:
        if l.startswith(".text.") or l.startswith("//-----"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            chain.append((m.group(1), int(m.group(2))))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            if chain:
                names = []
                for f, ln in reversed(chain):              # outermost first
                    nm = fn_of(f, ln)
                    if not names or names[-1] != nm:
                        names.append(nm)
                cur = tuple(names)
                chain = []
            paths.append(cur)
    k = min(len(paths), len(n))
    print("instructions: ncu %d, nvdisasm %d" % (len(n), len(paths)))
    tot = n[:k].sum()
    agg = defaultdict(float)
    mix = defaultdict(lambda: defaultdict(float))
    cnt = defaultdict(int)
    for i in range(k):
        key = paths[i][:depth]
        agg[key] += n[i]
        cnt[key] += 1
        for c in stalls:
            mix[key][c] += st[c][i]
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        m = sorted(mix[key].items(), key=lambda kv: -kv[1])[:3]
        print("%6.2f%%  %5d instr  %-70s %s" % (100 * v / tot, cnt[key], " > ".join(key)[-70:],
                                                 " ".join("%s %.0f%%" % (a.replace("stall_", ""), 100 * b / max(v, 1)) for a, b in m)))


if __name__ == "__main__":
    main()
