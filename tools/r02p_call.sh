#!/bin/bash
# Quick re-check of the dip / divne GPU tests after the CG head kernels moved to the fixed-piece mapping.
set -u
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "divne or dip3d_golden or dip2d_golden or dip3d_vs_oracle or dip3d_constant" > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc $?: $(tail -1 gpurun_out/r02p_pytest.log)" | tee gpurun_out/r02p_summary.txt
