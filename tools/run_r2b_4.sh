set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b4_bench.json 2> gpurun_out/r2b4_bench.err
tail -3 gpurun_out/r2b4_bench.err
BP_SCAN_UNIFORM=0 python bench.py --no-extra --cpu-budget 0.5 --steps 10 --warmup 3 > gpurun_out/r2b4_cfg5_general.json 2> gpurun_out/r2b4_cfg5_general.err
