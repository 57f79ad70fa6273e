#!/bin/bash
# 2-GPU call: slab parity with the register kernels of the axis-3 pass (128- and 32-plane slabs), timing on the slab
# geometry of 8 ranks (1000x1024x256 over 2 ranks), register kernels against the tile kernels.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/r02l_call.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
S=$O/r02l_summary_n$N.txt
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
echo "== $N GPUs: slab parity against the oracle and against a single-GPU run" | tee $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_check.py > $O/r02l_dist_check_n$N.log 2>&1; echo "dist_check rc $?" | tee -a $S
grep "dist_check" $O/r02l_dist_check_n$N.log | sort -u | tee -a $S
tail -2 $O/r02l_dist_check_n$N.log | tee -a $S
p=29540
for v in "PST_TRI3_SPLIT=1" "PST_TRI3_SPLIT=1 PST_TRI3_REG=0 PST_TRI3_WMAX=64"; do
    tag=$(echo "256_${v}" | tr ' =,' '___')
    p=$((p+1))
    env $v timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $p bench.py --gpus $N --shape 1000,1024,256 --steps 3 --warmup 3 --no-e2e > $O/r02l_bench_$tag.json 2> $O/r02l_bench_$tag.err
    echo "1000,1024,256 [$v] rc $?: $(line $O/r02l_bench_$tag.json)" | tee -a $S
done
