mkdir -p gpurun_out/r2f
python -m pytest tests -x -q -m gpu > gpurun_out/r2f/tests.log 2>&1; echo "tests rc=$?" 
tail -3 gpurun_out/r2f/tests.log
python bench.py > gpurun_out/r2f/bench_n1.json 2> gpurun_out/r2f/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/r2f/bench_n1.json
python bench.py --impl reference > gpurun_out/r2f/bench_ref.json 2> gpurun_out/r2f/bench_ref.err; cat gpurun_out/r2f/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f/launches.csv python bench.py --steps 1 --warmup 1 --no-legs --no-parity > gpurun_out/r2f/ncu_bench.log 2>&1; echo "ncu rc=$?"
