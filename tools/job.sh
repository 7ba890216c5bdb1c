python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/step_breakdown.py > gpurun_out/breakdown.txt 2>gpurun_out/breakdown.err
python bench.py --steps 5 --warmup 3 --profile-layers --no-cpu-baseline > gpurun_out/bench_r1_v6.json 2> gpurun_out/bench_r1_v6_layers.txt
tail -3 gpurun_out/bench_r1_v6_layers.txt
cut -c1-300 gpurun_out/bench_r1_v6.json
head -60 gpurun_out/breakdown.txt | cut -c1-140
