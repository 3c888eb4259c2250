# round 2, third session: the N = 2 tests of the frame with round splitters + a phase trace at 2 x 2^25
set -x
timeout 300 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r2c_gputests_n2.log
cat gpurun_out/r2c_gputests_n2.log
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_trace.py 25 > gpurun_out/r2c_trace_n2.log 2>&1
tail -12 gpurun_out/r2c_trace_n2.log
