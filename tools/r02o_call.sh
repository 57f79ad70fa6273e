#!/bin/bash
# Re-check of the multi-rank axis-3 kernels on one GPU after the per-line code moved into pst_tri3_reg_core.h.
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests -q -m gpu -k "axis3" > gpurun_out/r02o_pytest.log 2>&1; echo "pytest -k axis3 rc $?: $(tail -1 gpurun_out/r02o_pytest.log)" | tee gpurun_out/r02o_summary.txt
