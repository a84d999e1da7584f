mkdir -p gpurun_out/r2o
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests/test_search_gpu.py tests/test_align_gpu.py -x -q 2>&1 | tail -2
python bench.py --no-legs --no-job --steps 3 2>gpurun_out/r2o/b.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity_checked'), d.get('parity_mismatches'), d.get('walks_skipped_fraction'))"
for K in 'traceback_ckpt_tasks_kernel<\(int\)8>' 'nw_ckpt_kernel<\(int\)8, \(int\)0>'; do
  N=$(echo "$K" | tr -dc 'a-z_0-9')
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 1 -c 1 -f -o gpurun_out/r2o/$N python tools/stage_times.py 32768 --short > gpurun_out/r2o/ncu_$N.log 2>&1; echo "ncu $N rc=$?"
  python tools/ncu_summary.py gpurun_out/r2o/$N.ncu-rep > gpurun_out/r2o/${N}_summary.txt 2>&1
  head -4 gpurun_out/r2o/${N}_summary.txt
done
rm -f gpurun_out/r2o/nw_ckpt_kernelint8int0.ncu-rep
