mkdir -p gpurun_out/r2p
python -m pytest tests/test_group_gpu.py -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2p/bench_n2.json 2> gpurun_out/r2p/bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2p/bench_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2p/bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','parity_checked','parity_mismatches')}, d['e2e']['value'], {k:(v.get('value'),v.get('ms_per_step')) for k,v in d.get('legs',{}).items()}, d.get('job_mode'))
P
