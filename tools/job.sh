python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 8 --warmup 3 --profile-layers --no-cpu-baseline > gpurun_out/bench_r1_v9.json 2> gpurun_out/bench_r1_v9_layers.txt
python - <<'PY'
import json
for l in open('gpurun_out/bench_r1_v9.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d[k] for k in ['value','ms_per_step','e2e','eager_ms_per_step','gpu_launches_per_step']}); print(d['roofline']['achieved'], d['roofline']['all_tcgen05'], d['roofline']['hbm_bound_convs'])
PY
