set -x
python -m pytest tests/test_gpu_sort_finish.py -x -q -m gpu 2>&1 | tail -15
python -m pytest tests -q -m gpu --deselect tests/test_gpu_sort_finish.py 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err
tail -3 gpurun_out/r2b_bench1.err
