#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests/test_gpu_forms.py tests/test_gpu_guard.py -x -q -m gpu 2>&1 | tail -1
