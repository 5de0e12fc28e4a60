T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29593"
timeout 300 $T tools/cfg5_bench.py --reps 2 2> gpurun_out/cfg5_n8.err > gpurun_out/cfg5_n8.json; echo cfg5 rc=$?
grep "^{" gpurun_out/cfg5_n8.json
timeout 300 $T tools/dist_check.py > gpurun_out/dist_n8.log 2>&1; echo dist rc=$?
grep "PASS\|FAIL\|rror" gpurun_out/dist_n8.log | tail -6
