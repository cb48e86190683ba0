#!/bin/bash
# usage: tools/sweep.sh <tag> <bench args...> -- lib1.so lib2.so ...   (runs bench.py once per library variant)
tag=$1; shift
args=()
while [[ $# -gt 0 && "$1" != "--" ]]; do args+=("$1"); shift; done
shift
mkdir -p gpurun_out
for lib in "$@"; do
  name=$(basename $lib .so)
  UFAIR_LIB=$PWD/$lib python bench.py "${args[@]}" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${name}.json"))
    r = d.get("roofline", {})
    print("${tag} ${name}: kernel_ms %.3f step_ms %.3f frac %.4f variant %s clocks %s" % (d["kernel_ms_per_launch"], d["ms_per_step"], r.get("frac", 0), r.get("kernel_variant"), (d.get("clocks") or {}).get("sm_mhz")))
except Exception as e:
    print("${tag} ${name}: FAILED", e)
PY
done
