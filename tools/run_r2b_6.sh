set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sort_finish.py tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --cpu-budget 0.5 > gpurun_out/r2b6_bench.json 2> gpurun_out/r2b6_bench.err
tail -3 gpurun_out/r2b6_bench.err
