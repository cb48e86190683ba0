#!/bin/bash
mkdir -p gpurun_out
PYTHONPATH=$PWD python tools/experiments/scenario_e2e_timeline.py 262144 > gpurun_out/scen_timeline.txt 2>&1; head -70 gpurun_out/scen_timeline.txt
