timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu15.log
timeout 120 python tools/profile_run.py --frames 4 --reps 3 | tail -1
timeout 120 python tools/profile_run.py --frames 16 --reps 2 | tail -1
timeout 600 python bench.py > gpurun_out/bench16.log 2>&1; echo "bench rc=$?"; grep -o '"ref_passes_per_s": [0-9.]*' gpurun_out/bench16.log | head -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches16d.csv python tools/profile_run.py --frames 16 --reps 1 > /dev/null 2>&1
