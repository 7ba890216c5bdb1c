python tools/ncu_eval.py > gpurun_out/plain_eval.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_fused_kernel -s 4 -c 2 -o gpurun_out/prof_eval_r1_v10 python tools/ncu_eval.py > gpurun_out/ncu_eval.log 2>&1
tail -2 gpurun_out/ncu_eval.log
