#!/bin/bash
# round-2 GPU call 40: full GPU suite + smoke at the round's last commit
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_t40.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t40.log
tail -4 gpurun_out/r02_t40.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
