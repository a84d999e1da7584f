mkdir -p gpurun_out/r2g
python tools/stage_times.py 32768 --short > gpurun_out/r2g/stage_c2.log 2>&1; cat gpurun_out/r2g/stage_c2.log
python tools/stage_times.py 16384 --short --c4 > gpurun_out/r2g/stage_c4.log 2>&1; cat gpurun_out/r2g/stage_c4.log
for K in rank_kernel traceback_ckpt_tasks nw_ckpt_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/r2g/$K python tools/stage_times.py 32768 --short > gpurun_out/r2g/ncu_$K.log 2>&1; echo "ncu $K rc=$?"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rank_kernel -s 1 -c 1 -f -o gpurun_out/r2g/rank_c4 python tools/stage_times.py 16384 --short --c4 > gpurun_out/r2g/ncu_rank_c4.log 2>&1; echo "ncu rank c4 rc=$?"
ls -la gpurun_out/r2g
for K in rank_kernel traceback_ckpt_tasks nw_ckpt_kernel rank_c4; do
  if [ -f gpurun_out/r2g/$K.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/r2g/$K.ncu-rep > gpurun_out/r2g/${K}_summary.txt 2>&1
    ncu -i gpurun_out/r2g/$K.ncu-rep --page raw --csv > gpurun_out/r2g/${K}_raw.csv 2>/dev/null
  fi
done
rm -f gpurun_out/r2g/nw_ckpt_kernel.ncu-rep gpurun_out/r2g/rank_c4.ncu-rep
du -sh gpurun_out/r2g
