mkdir -p gpurun_out
for W in "64,32,64,2,32" "128,64,128,2,32"; do
ALG_WIDTHS=$W timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gen_launches.csv python tools/quick_bench.py 12 3 3 2097152 generic > gpurun_out/gen_prof.log 2>&1
tail -2 gpurun_out/gen_prof.log | head -1 | cut -c1-150
python - <<'PY'
import csv, re, collections
rows = list(csv.reader(l for l in open("gpurun_out/gen_launches.csv") if not l.startswith("==")))
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    k = re.sub(r"\(.*", "", r[ki])[:70]
    agg[k] += ms; cnt[k] += 1
tot = sum(agg.values())
print("total %.2f ms" % tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:16]:
    print("%8.3f ms %5.1f%% %5d  %s" % (v, 100 * v / tot, cnt[k], k))
PY
done
