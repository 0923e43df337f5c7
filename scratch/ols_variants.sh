#!/bin/bash
# usage: ols_variants.sh base v1 v2 ...   (timing part of check_ols.py for every variant library)
for v in "$@"; do
  if [ "$v" = "base" ]; then unset BASIC_DSP_B200_LIB; else export BASIC_DSP_B200_LIB=/root/repo/scratch/lib_$v.so; fi
  echo "== $v"
  python scratch/check_ols.py | grep "taps=\|FAIL"
done
