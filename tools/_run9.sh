mkdir -p gpurun_out/r2m
python tools/sweep_search.py 2>&1 | tee gpurun_out/r2m/sweep.log | grep subbatch
VSG_TRACE=1 VSG_HOST_THREADS=1 python tools/sweep_one.py 2>&1 | grep "vsg trace\] batch" | tail -4
