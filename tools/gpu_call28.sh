#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/tests28.txt 2>&1; tail -2 gpurun_out/tests28.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
