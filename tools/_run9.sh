#!/bin/bash
# development: timing of variant libraries at 1080p and 4K
for v in $VARIANTS; do
  if [ "$v" = main ]; then L=vvc-affine-gpu_b200/libaffine_me.so; else L=build_variants/libaffine_me_$v.so; fi
  echo "== $v"
  export AME_LIB=$PWD/$L
  python tools/profile_run.py --frames 16 --reps 3 | tail -1
  python tools/profile_run.py --frames 6 --reps 2 --size 3840x2160 | tail -1
done
