#!/bin/bash
for v in "$@"; do
  if [ "$v" = "base" ]; then unset BASIC_DSP_B200_LIB; else export BASIC_DSP_B200_LIB=/root/repo/scratch/lib_$v.so; fi
  echo "== $v"; BDSP_OLS_FORCE=4096 python scratch/ols_choice.py 31 1023
done
