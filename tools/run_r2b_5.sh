set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b5_bench.json 2> gpurun_out/r2b5_bench.err
tail -3 gpurun_out/r2b5_bench.err
python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2b5_profile_cfg5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"scan_groups|pair_finish" -s 2 -c 2 -o gpurun_out/r2b5_k -f \
    python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2b5_ncu.log 2>&1
ncu -i gpurun_out/r2b5_k.ncu-rep --page raw --csv > gpurun_out/r2b5_k.raw.csv 2>/dev/null
ncu -i gpurun_out/r2b5_k.ncu-rep --page source --csv > gpurun_out/r2b5_k.source.csv 2>/dev/null
rm -f gpurun_out/r2b5_k.ncu-rep
gzip -f gpurun_out/r2b5_k.source.csv
