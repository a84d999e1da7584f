mkdir -p gpurun_out/r2i
python -m pytest tests/test_search_gpu.py tests/test_udb_gpu.py tests/test_scale_gpu.py tests/test_mask_gpu.py -x -q > gpurun_out/r2i/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2i/tests.log
python tools/stage_times.py 32768 --short > gpurun_out/r2i/stage_c2.log 2>&1; cat gpurun_out/r2i/stage_c2.log
python tools/stage_times.py 16384 --short --c4 > gpurun_out/r2i/stage_c4.log 2>&1; cat gpurun_out/r2i/stage_c4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rank_kernel -s 1 -c 1 -f -o gpurun_out/r2i/rank_kernel python tools/stage_times.py 32768 --short > gpurun_out/r2i/ncu_rank.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r2i/rank_kernel.ncu-rep > gpurun_out/r2i/rank_kernel_summary.txt 2>&1; cat gpurun_out/r2i/rank_kernel_summary.txt
ncu -i gpurun_out/r2i/rank_kernel.ncu-rep --page raw --csv > gpurun_out/r2i/rank_kernel_raw.csv 2>/dev/null
