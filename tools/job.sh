python -m pytest tests/test_encoder_gpu.py tests/test_ops_gpu.py tests/test_modules_gpu.py -q 2>&1 | tail -4
python tools/step_breakdown.py > gpurun_out/breakdown.txt 2>gpurun_out/breakdown.err
head -36 gpurun_out/breakdown.txt | cut -c1-130
