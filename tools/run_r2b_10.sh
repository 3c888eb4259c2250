set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --cpu-budget 0.5 > gpurun_out/r2b10_bench.json 2> gpurun_out/r2b10_bench.err
tail -3 gpurun_out/r2b10_bench.err
