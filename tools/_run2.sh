python tools/stage_times.py 16384 --short 2>&1 | grep -E "rank " | tail -1
python -m pytest tests/test_search_gpu.py -x -q 2>&1 | tail -1
