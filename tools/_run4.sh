#!/bin/bash
# development: ncu launch lists (serialised, per-kernel times) of variant libraries
for v in $VARIANTS; do
  if [ "$v" = main ]; then L=vvc-affine-gpu_b200/libaffine_me.so; else L=build_variants/libaffine_me_$v.so; fi
  AME_LIB=$PWD/$L timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ll_$v.csv python tools/profile_run.py --frames 16 --reps 1 > /dev/null 2>&1
  echo "$v rc=$?"
done
