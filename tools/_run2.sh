mkdir -p gpurun_out/r2n2
python -m pytest tests/test_group_gpu.py -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/r2n2/bench_n2.json 2> gpurun_out/r2n2/bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2n2/bench_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2n2/bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], {k:(v.get('value'),v.get('ms_per_step')) for k,v in d.get('legs',{}).items()})
P
python bench.py --no-legs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n1', d['value'], d['ms_per_step'])"
