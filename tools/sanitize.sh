# compute-sanitizer passes over the newest kernels (sparse index build, flat posting loop, gated traceback); run under gpurun
mkdir -p gpurun_out/san
export VSG_HOST_THREADS=2
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/san/memcheck.log python -m pytest -x -q "tests/test_search_gpu.py::test_rank_and_search_golden" "tests/test_search_gpu.py::test_traceback_on_demand_does_not_change_the_hit_tables" "tests/test_udb_gpu.py::test_wordlength_above_10_vs_compiled_reference" "tests/test_search_gpu.py::test_long_queries_rank_vs_oracle" > gpurun_out/san/memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/san/memcheck_pytest.log; grep -c "Invalid\|Error" gpurun_out/san/memcheck.log; tail -5 gpurun_out/san/memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/san/racecheck.log python -m pytest -x -q "tests/test_search_gpu.py::test_rank_and_search_golden" > gpurun_out/san/racecheck_pytest.log 2>&1; echo "racecheck rc=$?"
tail -2 gpurun_out/san/racecheck_pytest.log; tail -6 gpurun_out/san/racecheck.log
