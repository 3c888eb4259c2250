set -x
python -m pytest tests/test_gpu_sort_finish.py -x -q -m gpu 2>&1 | tail -15
python -m pytest tests/test_gpu_parity.py tests/test_reference_fixtures.py -q -m gpu 2>&1 | tail -5
python bench.py --workload cfg3 --no-extra --cpu-budget 1 --steps 10 --warmup 3 > gpurun_out/r2b_cfg3_b.json 2> gpurun_out/r2b_cfg3_b.err
python bench.py --workload cfg5 --no-extra --cpu-budget 1 --steps 10 --warmup 3 > gpurun_out/r2b_cfg5_b.json 2> gpurun_out/r2b_cfg5_b.err
python tools/profile_frame.py cfg3 16777216 2 > gpurun_out/r2b_profile_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:record_finish -s 1 -c 1 -o gpurun_out/r2b_finish -f \
    python tools/profile_frame.py cfg3 16777216 2 > gpurun_out/r2b_ncu_finish.log 2>&1
ncu -i gpurun_out/r2b_finish.ncu-rep --page raw --csv > gpurun_out/r2b_finish.raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_finish.ncu-rep --page source --csv > gpurun_out/r2b_finish.source.csv 2>/dev/null
rm -f gpurun_out/r2b_finish.ncu-rep
gzip -f gpurun_out/r2b_finish.source.csv
tail -2 gpurun_out/r2b_ncu_finish.log
