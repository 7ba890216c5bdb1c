set -x
python tools/ncu_conv.py > gpurun_out/plain_conv.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 6 -c 3 -o gpurun_out/prof_conv_r1_v7 python tools/ncu_conv.py > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/ncu_conv.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eval > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 1300 --csv --log-file gpurun_out/launches_r1_v7.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eval > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_r1_v7.csv
