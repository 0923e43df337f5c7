#!/bin/bash
for v in "$@"; do
  if [ "$v" = "base" ]; then unset BASIC_DSP_B200_LIB; else export BASIC_DSP_B200_LIB=/root/repo/scratch/lib_$v.so; fi
  python bench_configs.py --iters 20 --configs C4a | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'ms %.4f' % d['ms_median'], 'frac %.3f' % d['roofline_frac'])"
done
