set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_final_bench_n8.json 2> gpurun_out/r2_final_bench_n8.err
tail -2 gpurun_out/r2_final_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/dist_trace.py 25 > gpurun_out/r2_final_trace_n8.log 2>&1
grep -A2 "^rank 0" gpurun_out/r2_final_trace_n8.log
