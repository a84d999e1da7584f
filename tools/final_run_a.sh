#!/bin/bash
# round-end validation on the GPU box: full GPU test suite, the bench arms, the secondary workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
timeout 300 python bench.py --workload allpairs > gpurun_out/bench_allpairs_n1.json 2> gpurun_out/bench_allpairs.err; tail -c 600 gpurun_out/bench_allpairs_n1.json
