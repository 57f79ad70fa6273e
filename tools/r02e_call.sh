#!/bin/bash
# Multi-GPU call (N = number of visible GPUs): slab parity against the oracle, then bench.py at N with parity_vs_n1.
#   gpurun --gpus 2 --timeout 1800 -- 'bash tools/r02e_call.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
S=$O/r02e_summary_n$N.txt
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
run() { # workload steps extra...
    w=$1; shift
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload $w "$@" > $O/r02e_bench_${w}_n$N.json 2> $O/r02e_bench_${w}_n$N.err
    echo "$w N=$N rc $?: $(line $O/r02e_bench_${w}_n$N.json)" | tee -a $S
}
echo "== $N GPUs: slab parity against the oracle (tests/dist_check.py at world = $N)" | tee $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_check.py > $O/r02e_dist_check_n$N.log 2>&1; echo "dist_check rc $?" | tee -a $S
grep "dist_check" $O/r02e_dist_check_n$N.log | tee -a $S
tail -3 $O/r02e_dist_check_n$N.log | tee -a $S
echo "== bench.py at N = $N" | tee -a $S
run dip3d_somf3d --steps 3 --warmup 3
PST_TRI3_PEERHALO=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 --no-e2e --no-parity > $O/r02e_bench_nccl_halo_n$N.json 2> $O/r02e_bench_nccl_halo_n$N.err
echo "headline, NCCL halos (PST_TRI3_PEERHALO=0) rc $?: $(line $O/r02e_bench_nccl_halo_n$N.json)" | tee -a $S
PST_CG_DEVSCALARS=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 2 --warmup 3 --no-e2e --no-parity > $O/r02e_bench_hostscalars_n$N.json 2> $O/r02e_bench_hostscalars_n$N.err
echo "headline, host-side CG scalars (PST_CG_DEVSCALARS=0) rc $?: $(line $O/r02e_bench_hostscalars_n$N.json)" | tee -a $S
run soint3d --steps 3 --warmup 3
run sint3d --steps 2 --warmup 2
if [ $N -le 2 ]; then run somean3d --steps 5 --warmup 3; fi
