set -x
python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_final_gputests_n2.log
cat gpurun_out/r2_final_gputests_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_final_bench_n2.json 2> gpurun_out/r2_final_bench_n2.err
tail -2 gpurun_out/r2_final_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_trace.py 25 > gpurun_out/r2_final_trace_n2.log 2>&1
tail -12 gpurun_out/r2_final_trace_n2.log
