#!/bin/bash
# tools/variants2.sh "THREADS SLOTS MIN_CTAS SLACK" ...  -> build/var/var_T_S_C_K.so
set -e
mkdir -p /root/repo/build/var
cd /root/repo/dynast-release_b200/csrc
for v in "$@"; do
  set -- $v
  name=var_$1_$2_$3_$4
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DDYN_TALLY_THREADS=$1 -DDYN_TALLY_SLOTS=$2 -DDYN_TALLY_MIN_CTAS=$3 -DDYN_TALLY_SLACK=$4 $EXTRA -c tally.cu -o /root/repo/build/var/$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/build/var/$name.so /root/repo/build/var/$name.o $(ls *.o | grep -v '^tally.o$') -lz -lpthread
  echo built $name
done
