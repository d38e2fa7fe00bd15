#!/bin/bash
# Build scratch/libsng_<name>.so with extra -D flags for the float32 engine (tuning sweeps; scratch/ is git-ignored
# but travels to the GPU box).  Usage: scripts/build_variant.sh name -DSNG_X_FOO=1 ...
set -e
name=$1; shift
cd "$(dirname "$0")/../smart_nanogrid_gym_b200/csrc"
mkdir -p ../../scratch
make -s ../libsng.so >/dev/null
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC "$@" -c -o ../../scratch/sng_api_$name.o sng_api.cu
nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../../scratch/libsng_$name.so ../../scratch/sng_api_$name.o sng_f64.o sng_policy.o sng_policy_tc.o -lcudart
rm -f ../../scratch/sng_api_$name.o
echo built scratch/libsng_$name.so
