python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu8.log
python tools/profile_run.py --frames 4 --reps 3 > gpurun_out/pp10.log 2>&1; cat gpurun_out/pp10.log
python tools/profile_run.py --frames 16 --reps 2 > gpurun_out/pp11.log 2>&1; cat gpurun_out/pp11.log
ncu --set full --clock-control none --import-source on -k regex:ame_iter_small --launch-skip 2 --launch-count 1 -o gpurun_out/prof_one16 -f python tools/profile_run.py --frames 16 --reps 1 > gpurun_out/ncu_one16.log 2>&1; echo "ncu rc=$?"
python bench.py > gpurun_out/bench10.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench10.log | cut -c1-900
