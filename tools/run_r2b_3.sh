set -x
python -m pytest tests/test_gpu_sort_finish.py -x -q -m gpu 2>&1 | tail -5
BP_SORT_FINISH_WALK=1 python -m pytest tests/test_gpu_sort_finish.py -x -q -m gpu 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "config3 or full_size" 2>&1 | tail -3
B="python bench.py --no-extra --cpu-budget 0.5 --steps 10 --warmup 3"
$B --workload cfg3 > gpurun_out/r2b3_cfg3.json 2> gpurun_out/r2b3_cfg3.err
$B --workload cfg4 > gpurun_out/r2b3_cfg4.json 2> gpurun_out/r2b3_cfg4.err
BP_PAIR_DROP_MAX_GROUP=16 $B --workload cfg4 > gpurun_out/r2b3_cfg4_drop16.json 2> gpurun_out/r2b3_cfg4_drop16.err
BP_RADIX_BITS=9 $B --workload cfg4 > gpurun_out/r2b3_cfg4_rb9.json 2> gpurun_out/r2b3_cfg4_rb9.err
BP_RADIX_BITS=9 $B --workload cfg5 > gpurun_out/r2b3_cfg5_rb9.json 2> gpurun_out/r2b3_cfg5_rb9.err
python tools/profile_frame.py cfg3 16777216 2 > gpurun_out/r2b_profile_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:record_finish -s 1 -c 1 -o gpurun_out/r2b_finish -f \
    python tools/profile_frame.py cfg3 16777216 2 > gpurun_out/r2b_ncu_finish.log 2>&1
ncu -i gpurun_out/r2b_finish.ncu-rep --page raw --csv > gpurun_out/r2b_finish.raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_finish.ncu-rep --page source --csv > gpurun_out/r2b_finish.source.csv 2>/dev/null
rm -f gpurun_out/r2b_finish.ncu-rep
gzip -f gpurun_out/r2b_finish.source.csv
