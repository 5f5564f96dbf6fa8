#!/bin/bash
# Runs on the GPU box (gpurun -- 'bash tools/collect_evidence.sh'): tests, bench lines, ncu launch lists and full
# captures of the four search kernels.  Everything lands in gpurun_out/ev_*; tools/summarize_evidence.py turns it
# into the tracked summaries under profiles/.  Numbers printed under ncu are never bench values.
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/ev_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/ev_pytest_gpu.log
timeout 900 python bench.py > $O/ev_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/ev_bench_reference.log 2>&1; echo "ref rc=$?"
timeout 300 python tools/profile_run.py --frames 16 --reps 3 > $O/ev_profile_run16.log 2>&1
timeout 300 python tools/profile_run.py --frames 4 --reps 3 > $O/ev_profile_run4.log 2>&1
# launch list of the bench command itself (first 600 launches: untimed uploads + launch sequences of the warm-up steps)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ev_launches_bench.csv python bench.py --steps 1 --warmup 3 > $O/ev_ncu_bench.log 2>&1
# launch list + DRAM bytes of one complete launch sequence (16 frames = 58 searches)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ev_launches_seq16.csv python tools/profile_run.py --frames 16 --reps 1 > /dev/null 2>&1
# full captures (source-level) of the search kernels, third iteration of the 2-CP search
for k in ame_iter_small ame_iter_big ame_update_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 --launch-count 1 -o $O/ev_$k -f python tools/profile_run.py --frames 16 --reps 1 > $O/ev_ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ame_iter0_kernel --launch-count 1 -o $O/ev_ame_iter0_kernel -f python tools/profile_run.py --frames 16 --reps 1 > $O/ev_ncu_iter0.log 2>&1
echo done
