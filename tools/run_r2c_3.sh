# round 2, third session: phase trace + result check of the 8 x 2^25 frame (config 5) with round splitters
set -x
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/dist_trace.py 25 > gpurun_out/r2c_trace_n8.log 2>&1
grep -v "^\*\|OMP_NUM" gpurun_out/r2c_trace_n8.log | tail -40
