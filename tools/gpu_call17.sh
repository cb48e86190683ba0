#!/bin/bash
mkdir -p gpurun_out
PYTHONPATH=$PWD python tools/experiments/scenario_e2e_chunks.py 2>&1 | tee gpurun_out/scen_chunks.txt
