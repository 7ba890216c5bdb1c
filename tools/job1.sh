set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python tools/step_breakdown.py > gpurun_out/breakdown.txt 2> gpurun_out/breakdown.err
python bench.py --steps 5 --warmup 3 --profile-layers > gpurun_out/bench_r1_v3.json 2> gpurun_out/bench_r1_v3_layers.txt
tail -3 gpurun_out/pytest_gpu.log; head -40 gpurun_out/breakdown.txt; tail -5 gpurun_out/breakdown.err
