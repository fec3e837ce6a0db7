#!/bin/bash
# build tally variants: tools/variants.sh "T S C" ...   -> build/var_T_S_C.so
set -e
mkdir -p build/var
cd /root/repo/dynast-release_b200/csrc
for v in "$@"; do
  set -- $v
  name=var_$1_$2_$3
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DDYN_TALLY_THREADS=$1 -DDYN_TALLY_SLOTS=$2 -DDYN_TALLY_MIN_CTAS=$3 $EXTRA -c tally.cu -o /root/repo/build/var/$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/build/var/$name.so /root/repo/build/var/$name.o $(ls *.o | grep -v '^tally.o$') -lz -lpthread
  echo built $name
done
