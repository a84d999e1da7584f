mkdir -p gpurun_out/tb2
for r in 0; do echo "refill=$r"; VSG_TB_REFILL=$r python tools/stage_times.py 16384 --short 2>&1 | tail -1; done
python -m pytest tests/test_align_gpu.py tests/test_stress_gpu.py tests/test_scale_gpu.py -x -q 2>&1 | tail -2
