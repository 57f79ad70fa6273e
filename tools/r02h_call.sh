#!/bin/bash
# N = 4 or 8 GPUs: slab parity against the oracle, bench.py at N with parity_vs_n1 (headline + the C4 interpolators).
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/r02h_call.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
S=$O/r02h_summary_n$N.txt
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 1), "| parity_vs_n1", d.get("parity_vs_n1"), (d.get("parity_detail") or {}).get("rel_l2"), "bit_exact", (d.get("parity_detail") or {}).get("bit_exact"),
          "| e2e", round(d.get("e2e", {}).get("value", 0), 2), {k: round(v["ms_per_step"], 1) for k, v in r.get("classes", {}).items()})
except Exception as e:
    print("no JSON line:", e)
PY
}
echo "== $N GPUs: slab parity against the oracle (tests/dist_check.py at world = $N)" | tee $S
if [ -z "${SKIP_CHECK:-}" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_check.py > $O/r02h_dist_check_n$N.log 2>&1; echo "dist_check rc $?" | tee -a $S
grep "dist_check" $O/r02h_dist_check_n$N.log | sort -u | tee -a $S
fi
echo "== bench.py at N = $N" | tee -a $S
p=29560
for w in ${WORKLOADS:-dip3d_somf3d soint3d sint3d}; do
    p=$((p+1))
    steps=3; [ $w = sint3d ] && steps=2
    PST_TRI3_SPLIT=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $p bench.py --gpus $N --workload $w --steps $steps --warmup 3 > $O/r02h_bench_${w}_n$N.json 2> $O/r02h_bench_${w}_n$N.err
    echo "$w N=$N rc $?: $(line $O/r02h_bench_${w}_n$N.json)" | tee -a $S
done
