#!/usr/bin/env python
"""Warp-stall sampling by SASS instruction from an `ncu --set full --import-source on` report:
   python tools/ncu_stalls.py <report.ncu-rep> <kernel-regex> [top=30]
prints the stall-reason mix, the hottest instructions and the per-opcode share (development aid;
the mbarrier try_wait spin loops show up as `@!P0 BRA` with stall_long_sb)."""
import io
import subprocess
import sys

import pandas as pd


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout.split("\n")
    hdr = [i for i, l in enumerate(raw) if l.startswith('"Address"')]
    for hi, start in enumerate(hdr):
        end = hdr[hi + 1] - 1 if hi + 1 < len(hdr) else len(raw)
        print("==", raw[start - 1][:120])
        df = pd.read_csv(io.StringIO("\n".join(raw[start:end])))
        df["n"] = pd.to_numeric(df["# Samples"], errors="coerce").fillna(0)
        tot = df["n"].sum()
        stalls = [c for c in df.columns if c.startswith("stall_") and "Not Issued" not in c]
        for c in stalls:
            df[c] = pd.to_numeric(df[c], errors="coerce").fillna(0)
        mix = (df[stalls].sum() / tot).sort_values(ascending=False).head(8)
        print("samples %d | " % tot + "  ".join("%s %.1f%%" % (k.replace("stall_", ""), 100 * v) for k, v in mix.items()))
        for _, r in df.sort_values("n", ascending=False).head(top).iterrows():
            st = max(stalls, key=lambda c: r[c])
            print("%7d %5.2f%%  %-64s %s" % (r["n"], 100 * r["n"] / tot, r["Source"].strip()[:64], st.replace("stall_", "")))
        op = df["Source"].str.strip().str.replace(r"^@!?U?P\d+\s+", "", regex=True).str.split().str[0]
        print((100 * df.groupby(op)["n"].sum().sort_values(ascending=False).head(16) / tot).round(1).to_dict())


if __name__ == "__main__":
    main()
