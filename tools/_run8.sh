#!/bin/bash
# development: tests, then timing at the three frame sizes
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu23.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu23.log
python tools/profile_run.py --frames 16 --reps 3 | tail -1
python tools/profile_run.py --frames 8 --reps 2 --size 3840x2160 | tail -1
python tools/profile_run.py --frames 4 --reps 2 --size 7680x4320 | tail -1
