#!/bin/bash
# Final single-GPU verification of round 2: smoke + the whole GPU suite on the defaults.
set -u
mkdir -p gpurun_out
O=gpurun_out
S=$O/r02m_summary.txt
echo "== smoke + full GPU suite" | tee $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02m_smoke.log 2>&1; echo "smoke rc $?: $(tail -1 $O/r02m_smoke.log)" | tee -a $S
timeout 900 python -m pytest tests -q -m gpu --durations=5 -s > $O/r02m_pytest.log 2>&1; echo "pytest -m gpu rc $?" | tee -a $S
grep -E "passed|failed|rel-L2|panel\]|Error" $O/r02m_pytest.log | tail -8 | tee -a $S
