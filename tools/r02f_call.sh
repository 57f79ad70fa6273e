#!/bin/bash
# 2-GPU call: parity of the reworked distributed axis-3 kernels, then A/B timings of the N = 4 / N = 8 slab geometries
# emulated with two ranks (256- and 128-plane slabs).  gpurun --gpus 2 --timeout 1500 -- 'bash tools/r02f_call.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
S=$O/r02f_summary_n$N.txt
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 1), "| parity_vs_n1", d.get("parity_vs_n1"), (d.get("parity_detail") or {}).get("rel_l2"),
          "| e2e", round(d.get("e2e", {}).get("value", 0), 2), {k: round(v["ms_per_step"], 1) for k, v in r.get("classes", {}).items()})
except Exception as e:
    print("no JSON line:", e)
PY
}
echo "== $N GPUs: slab parity against the oracle" | tee $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_check.py > $O/r02f_dist_check_n$N.log 2>&1; echo "dist_check rc $?" | tee -a $S
grep "dist_check" $O/r02f_dist_check_n$N.log | sort -u | tee -a $S
tail -2 $O/r02f_dist_check_n$N.log | tee -a $S
PST_TRI3_RC=0 PST_TRI3_WIN=0 PST_TRI3_PEERHALO=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tests/dist_check.py > $O/r02f_dist_check_old_n$N.log 2>&1; echo "dist_check (round-1 kernels: PST_TRI3_RC=0 PST_TRI3_WIN=0 PST_TRI3_PEERHALO=0) rc $?: $(grep -c 'bit-exact=True\|rel-L2 0.00e+00' $O/r02f_dist_check_old_n$N.log) lines ok" | tee -a $S
p=29540
for shape in 1000,1024,1024 1000,1024,512 1000,1024,256; do
  for v in "" "PST_TRI3_RC=0" "PST_TRI3_RC=0 PST_TRI3_PEERHALO=0 PST_TRI3_WIN=0 PST_CG_DEVSCALARS=0"; do
    tag=$(echo "${shape}_${v:-default}" | tr ' =,' '___')
    p=$((p+1))
    extra="--no-e2e"; [ -z "$v" ] && extra=""
    env $v timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $p bench.py --gpus $N --shape $shape --steps 3 --warmup 3 $extra > $O/r02f_bench_$tag.json 2> $O/r02f_bench_$tag.err
    echo "$shape [$v] rc $?: $(line $O/r02f_bench_$tag.json)" | tee -a $S
  done
done
