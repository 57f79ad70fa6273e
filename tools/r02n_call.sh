#!/bin/bash
# Launch list of one headline step at 500x512x512 on the final defaults of round 2 (ncu, per-launch durations; serialised and
# cold-cache: shares, not absolute times).
set -u
mkdir -p gpurun_out
O=gpurun_out
S=$O/r02n_summary.txt
echo "== ncu: launch list of one headline step at 500x512x512" | tee $S
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02n_launches.csv python bench.py --shape 500,512,512 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/r02n_launches.log 2>&1; echo "launch list rc $?, $(wc -l < $O/r02n_launches.csv) lines" | tee -a $S
python - <<'PY' | tee -a $S
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02n_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
t = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    k = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    t[k] += float(r[vi].replace(",", "")); n[k] += 1
tot = sum(t.values())
for k, v in t.most_common(20):
    print(f"  {k:58s} {n[k]:6d} launches {v/1e6:9.2f} ms {100*v/tot:5.1f} %")
print(f"  total {tot/1e6:.1f} ms in {sum(n.values())} launches")
PY
gzip -f $O/r02n_launches.csv
