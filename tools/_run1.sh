timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu21.log
timeout 120 python tools/profile_run.py --frames 16 --reps 2 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_u3.csv -k regex:ame_update_kernel python tools/profile_run.py --frames 16 --reps 1 > /dev/null 2>&1
