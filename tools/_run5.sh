#!/bin/bash
# development: carve-out sweep
run() { echo "== $1 big=$2 small=$3"; AME_LIB=$PWD/$1 AME_CARVE_BIG=$2 AME_CARVE_SMALL=$3 timeout 200 python tools/profile_run.py --frames 16 --reps 3 | tail -1; }
M=vvc-affine-gpu_b200/libaffine_me.so
echo "== nored default"; AME_LIB=$PWD/build_variants/libaffine_me_nored.so timeout 200 python tools/profile_run.py --frames 16 --reps 3 | tail -1
echo "== main default"; AME_LIB=$PWD/$M timeout 200 python tools/profile_run.py --frames 16 --reps 3 | tail -1
run $M 50 -1
run $M 58 -1
run $M 72 -1
run $M 58 58
run $M 58 72
run $M 58 86
run $M 58 100
