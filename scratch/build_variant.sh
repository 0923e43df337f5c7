#!/bin/bash
# usage: build_variant.sh <name> <extra nvcc -D flags...>   -> scratch/lib_<name>.so
# (the flags go to the kernel files under study; everything else is taken from the regular build)
set -e
name=$1; shift
cd /root/repo/basic_dsp_b200/csrc
objs=""
for f in common fft conv ols4096i ols8192i ols64 fftp fftp16k fftc interp elementwise mathops reduce capi; do
  if [ "$f" = "ols4096i" ] || [ "$f" = "ols8192i" ] || [ "$f" = "fftp" ] || [ "$f" = "fftp16k" ] || [ "$f" = "conv" ] || [ "$f" = "fft" ] || [ "$f" = "interp" ] || [ "$f" = "ols64" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c $f.cu -o /tmp/var_${name}_$f.o &
    objs="$objs /tmp/var_${name}_$f.o"
  else
    objs="$objs ../build/$f.o"
  fi
done
wait
nvcc -shared -o /root/repo/scratch/lib_$name.so $objs -cudart static -Xlinker --exclude-libs,ALL
echo built scratch/lib_$name.so
