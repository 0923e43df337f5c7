#!/bin/bash
# usage: f64_variants.sh base c4 p10 ...   (variants built by scratch/build_variant.sh)
for v in "$@"; do
  if [ "$v" = "base" ]; then unset BASIC_DSP_B200_LIB; else export BASIC_DSP_B200_LIB=/root/repo/scratch/lib_$v.so; fi
  echo "== $v"
  python bench_configs.py --iters 10 --configs C5a | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C5a ms %.4f' % d['ms_median'], 'frac %.3f' % d['roofline_frac'])"
  python scratch/bench_fft_rows64.py | grep -E "n= +(16384|262144|1048576|4194304) "
done
