# Round-end evidence on one B200 (run under gpurun): GPU tests, default bench, reference arm, launch list, ncu --set full of the three kernels.
mkdir -p gpurun_out/round
( time python -m pytest tests -q -m gpu -x ) > gpurun_out/round/tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/round/tests.log
( time python bench.py > gpurun_out/round/bench_n1.json 2> gpurun_out/round/bench_n1.err ) 2> gpurun_out/round/bench_time.txt; echo "bench rc=$?"; tail -3 gpurun_out/round/bench_time.txt
python - <<'P'
import json
d=json.load(open('gpurun_out/round/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','parity_checked','parity_mismatches')}, d['e2e'], d['roofline']['frac'], d['roofline']['alone_ms'], d['clocks'])
print({k:(v.get('value'), v.get('ms_per_step'), v.get('e2e',{}).get('value') if isinstance(v.get('e2e'),dict) else None) for k,v in d.get('legs',{}).items()})
for k in ('lazy_mode','dust_mode','iupac_mode','job','cpu_baseline'):
    print(k, d.get(k))
P
( time python bench.py --impl reference > gpurun_out/round/bench_ref.json 2> gpurun_out/round/bench_ref.err ) 2>> gpurun_out/round/bench_time.txt; cat gpurun_out/round/bench_ref.json | cut -c1-600
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/round/launches.csv python bench.py --steps 1 --warmup 1 --no-legs --no-parity --no-cpu-baseline --no-job > gpurun_out/round/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
for K in 'traceback_ckpt_tasks_kernel<\(int\)8>' 'nw_ckpt_kernel<\(int\)8, \(int\)0>' 'rank_kernel'; do
  N=$(echo "$K" | tr -dc 'a-z_0-9')
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 1 -c 1 -f -o gpurun_out/round/$N python tools/stage_times.py 32768 --short > gpurun_out/round/ncu_$N.log 2>&1; echo "ncu $N rc=$?"
  python tools/ncu_summary.py gpurun_out/round/$N.ncu-rep > gpurun_out/round/${N}_summary.txt 2>&1
  ncu -i gpurun_out/round/$N.ncu-rep --page raw --csv > gpurun_out/round/${N}_raw.csv 2>/dev/null
  head -3 gpurun_out/round/${N}_summary.txt
done
du -sh gpurun_out/round
