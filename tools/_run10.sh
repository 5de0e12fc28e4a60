python tools/perf_modes.py 0,1 1,2,3 2>&1 | tail -6
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
