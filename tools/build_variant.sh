#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>: a copy of the library with tally.cu compiled under other macros
set -e
name=$1; shift
mkdir -p build/variants
cd dynast-release_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 "$@" -c tally.cu -o ../../build/variants/tally_$name.o
objs=$(ls *.o | grep -v '^tally.o$')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/variants/lib_$name.so $objs ../../build/variants/tally_$name.o -lz -lpthread
echo built build/variants/lib_$name.so
