python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --profile-layers > gpurun_out/bench_r1_v8.json 2> gpurun_out/bench_r1_v8_layers.txt
tail -2 gpurun_out/bench_r1_v8_layers.txt
python - <<'PY'
import json
for l in open('gpurun_out/bench_r1_v8.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d[k] for k in ['value','ms_per_step','e2e','eager_ms_per_step','clocks','gpu_launches_per_step']}); print(json.dumps(d['roofline'])[:1500]); print(d['eval']); print(d['cpu_baseline'])
PY
python tools/step_breakdown.py > gpurun_out/breakdown.txt 2>gpurun_out/breakdown.err
head -24 gpurun_out/breakdown.txt | cut -c1-120
