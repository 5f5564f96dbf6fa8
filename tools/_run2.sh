#!/bin/bash
# development A/B run on the GPU box: tests with the main library, then timing of every variant library
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu22.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu22.log
for v in "" $VARIANTS; do
  if [ -z "$v" ]; then L=vvc-affine-gpu_b200/libaffine_me.so; else L=build_variants/libaffine_me_$v.so; fi
  echo "== ${v:-main}"
  AME_LIB=$PWD/$L timeout 200 python tools/profile_run.py --frames 16 --reps 3 | tail -2
done
