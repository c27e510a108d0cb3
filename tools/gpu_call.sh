#!/bin/bash
# per-call scratch script (GPU box): last check of the final tree
mkdir -p gpurun_out
L=gpurun_out/call27.log
: > $L
echo "== pytest gpu" >> $L
timeout 100 python -m pytest tests -x -q -m gpu >> $L 2>&1
echo "rc=$?" >> $L
echo "== smoke" >> $L
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "rc=$?" >> $L
