timeout 900 python -m pytest tests -m gpu -x -q -k "share_first or seeded or reuse" > gpurun_out/pytest_gpu19.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu19.log
timeout 120 python tools/profile_run.py --frames 4 --reps 3 | tail -1
timeout 120 python tools/profile_run.py --frames 16 --reps 2 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches16h.csv -k regex:ame_iter0 python tools/profile_run.py --frames 16 --reps 1 > /dev/null 2>&1; grep iter0 gpurun_out/launches16h.csv | tail -1
