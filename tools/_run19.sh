python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/gpu_check.py perf 2>&1 | grep "infonce\|store\|fwd" | tail -4
python tools/itc_profile.py 2>&1 | grep "us/step" 
