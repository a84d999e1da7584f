mkdir -p gpurun_out/r2h
python -m pytest tests/test_udb_gpu.py -x -q > gpurun_out/r2h/udb.log 2>&1; echo "udb rc=$?"; tail -25 gpurun_out/r2h/udb.log
python -m pytest tests -q -m gpu --durations=25 --deselect tests/test_udb_gpu.py > gpurun_out/r2h/tests.log 2>&1; echo "tests rc=$?"; tail -45 gpurun_out/r2h/tests.log
