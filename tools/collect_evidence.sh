#!/bin/bash
# Runs on the GPU box (gpurun -- 'bash tools/collect_evidence.sh [quick]'): tests, bench lines of every configuration, ncu launch
# lists and full captures of the search kernels.  Everything lands in gpurun_out/ev_*; tools/summarize_evidence.py turns it
# into the tracked summaries under profiles/ (incl. profiles/r02_inst_table.json, which bench.py reads).  Numbers printed
# under ncu are never bench values.
set -x
O=gpurun_out
QUICK=${1:-full}
git rev-parse HEAD > $O/ev_head.txt 2>/dev/null || true
timeout 900 python -m pytest tests -m gpu -q > $O/ev_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/ev_pytest_gpu.log
timeout 900 python bench.py > $O/ev_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/ev_bench_reference.log 2>&1; echo "ref rc=$?"
if [ "$QUICK" != "quick" ]; then
  timeout 1200 python bench.py --config 4k --steps 2 --warmup 1 > $O/ev_bench_4k.log 2>&1; echo "4k rc=$?"
  timeout 1200 python bench.py --config 8k --steps 2 --warmup 1 > $O/ev_bench_8k.log 2>&1; echo "8k rc=$?"
fi
timeout 300 python tools/profile_run.py --frames 16 --reps 3 > $O/ev_profile_run16.log 2>&1
# launch list of the bench command itself (first 700 launches: plane preparation + launch sequences of the warm-up steps)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/ev_launches_bench.csv python bench.py --steps 1 --warmup 3 > $O/ev_ncu_bench.log 2>&1
# launch list + DRAM bytes + executed instructions of ONE launch sequence of the bench workload (64 frames = 250 searches, QP 32)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/ev_launches_seq64.csv python tools/profile_run.py --frames 64 --reps 1 > $O/ev_ncu_seq64.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/ev_launches_seq16.csv python tools/profile_run.py --frames 16 --reps 1 > /dev/null 2>&1
if [ "$QUICK" != "quick" ]; then
  timeout 1200 ncu --metrics $M --clock-control none --csv --log-file $O/ev_launches_seq64_4k.csv python tools/profile_run.py --size 3840x2160 --frames 64 --reps 1 > $O/ev_ncu_seq64_4k.log 2>&1
  timeout 1200 ncu --metrics $M --clock-control none --csv --log-file $O/ev_launches_seq16_8k.csv python tools/profile_run.py --size 7680x4320 --frames 16 --reps 1 > $O/ev_ncu_seq16_8k.log 2>&1
fi
# full captures (source-level) of the search kernels, third iteration of the 2-CP search (second of the 3-CP search for update<3>)
for k in ame_iter_small ame_iter_big ame_emit_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 --launch-count 1 -o $O/ev_$k -f python tools/profile_run.py --frames 16 --reps 1 > $O/ev_ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ame_update_kernel --launch-skip 2 --launch-count 1 -o $O/ev_ame_update_kernel2 -f python tools/profile_run.py --frames 16 --reps 1 > $O/ev_ncu_upd2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ame_update_kernel --launch-skip 7 --launch-count 1 -o $O/ev_ame_update_kernel3 -f python tools/profile_run.py --frames 16 --reps 1 > $O/ev_ncu_upd3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ame_iter0_kernel --launch-count 1 -o $O/ev_ame_iter0_kernel -f python tools/profile_run.py --frames 16 --reps 1 > $O/ev_ncu_iter0.log 2>&1
echo done
