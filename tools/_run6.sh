mkdir -p gpurun_out/r2j
python -m pytest tests/test_search_gpu.py tests/test_udb_gpu.py tests/test_scale_gpu.py -x -q > gpurun_out/r2j/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2j/tests.log
for V in "" tools/_libvsg_u4.so tools/_libvsg_u8.so; do
  echo "== lib=$V"
  VSG_LIB=$V python tools/stage_times.py 32768 --short 2>&1 | grep "^rank" | tail -1
  VSG_LIB=$V python tools/stage_times.py 16384 --short --c4 2>&1 | grep "^rank" | tail -1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rank_kernel -s 1 -c 1 -f -o gpurun_out/r2j/rank_kernel python tools/stage_times.py 32768 --short > gpurun_out/r2j/ncu_rank.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r2j/rank_kernel.ncu-rep > gpurun_out/r2j/rank_kernel_summary.txt 2>&1; cat gpurun_out/r2j/rank_kernel_summary.txt
