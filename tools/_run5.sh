python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/r10_pytest.log
python tools/stream_bench.py > gpurun_out/r10_stream.log 2>&1; echo rc=$?; cat gpurun_out/r10_stream.log | tail -12
