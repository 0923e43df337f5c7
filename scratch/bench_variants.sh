#!/bin/bash
for v in "$@"; do
  if [ "$v" = "base" ]; then unset BASIC_DSP_B200_LIB; else export BASIC_DSP_B200_LIB=/root/repo/scratch/lib_$v.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'ms/step %.4f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.0f' % d['e2e']['value'])"
done
