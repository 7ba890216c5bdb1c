for f in "--no-side-wgrad" "--no-side-wgrad" "" ""; do
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eval $f > gpurun_out/b.json 2> gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
for l in open('gpurun_out/b.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print("$f", {k:d[k] for k in ['value','ms_per_step','loss']})
PY
done
