for v in uloc usm; do echo "variant $v"; AME_LIB=$PWD/build_variants/libaffine_me_$v.so timeout 120 python tools/profile_run.py --frames 16 --reps 2 | tail -1
AME_LIB=$PWD/build_variants/libaffine_me_$v.so ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$v.csv -k regex:ame_update_kernel python tools/profile_run.py --frames 16 --reps 1 > /dev/null 2>&1
done
