mkdir -p gpurun_out/r2l
python -m pytest tests/test_search_gpu.py tests/test_scale_gpu.py tests/test_stream_gpu.py tests/test_seam2_gpu.py tests/test_group_gpu.py tests/test_dropin_gpu.py tests/test_align_gpu.py -x -q > gpurun_out/r2l/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2l/tests.log
VSG_TB_GATE=0 python bench.py --no-legs --no-cpu-baseline --no-job --steps 3 2>gpurun_out/r2l/b0.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gate0', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity_checked'), d.get('parity_mismatches'), d['roofline']['alone_ms'])"
python bench.py --no-legs --no-cpu-baseline --no-job --steps 3 2>gpurun_out/r2l/b1.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gate1', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity_checked'), d.get('parity_mismatches'), d['roofline']['alone_ms'])"
tail -3 gpurun_out/r2l/b1.err
