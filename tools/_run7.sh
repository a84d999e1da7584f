mkdir -p gpurun_out/r2k
python -m pytest tests/test_search_gpu.py tests/test_udb_gpu.py -x -q > gpurun_out/r2k/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2k/tests.log
for V in "" tools/_libvsg_u4z0.so tools/_libvsg_u6z1.so; do
  echo "== lib=$V"
  VSG_LIB=$V python tools/stage_times.py 32768 --short 2>&1 | grep "^rank" | tail -1
  VSG_LIB=$V python tools/stage_times.py 16384 --short --c4 2>&1 | grep "^rank" | tail -1
done
