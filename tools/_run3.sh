#!/bin/bash
# development A/B run on the GPU box: timing of every variant library (no tests)
for v in "" $VARIANTS; do
  if [ -z "$v" ]; then L=vvc-affine-gpu_b200/libaffine_me.so; else L=build_variants/libaffine_me_$v.so; fi
  echo "== ${v:-main}"
  AME_LIB=$PWD/$L timeout 200 python tools/profile_run.py --frames 16 --reps 3 | tail -2
done
