#!/bin/bash
# usage: tools/build_variant.sh <name> <extra nvcc flags...>  ->  fiveeqscm_b200/libufair_<name>.so
# (perf experiments: the <double, 3 gases> integrator with the given -D switches, everything else stubbed)
set -e
name=$1; shift
cd "$(dirname "$0")/../fiveeqscm_b200/csrc"
B=build_$name; mkdir -p $B
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC $*"
for f in ufair_abi ufair_host ufair_sampler ufair_inst_f64_g3; do nvcc $FLAGS -c $f.cu -o $B/$f.o & done
nvcc $FLAGS -c ../../tools/exp_stubs.cu -o $B/exp_stubs.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libufair_$name.so $B/*.o -lcudart
echo built libufair_$name.so
