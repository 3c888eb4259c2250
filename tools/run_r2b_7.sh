set -x
python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2b7_profile_cfg5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"scan_groups|pair_finish|encode_kernel" -s 3 -c 3 -o gpurun_out/r2b7_k -f \
    python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2b7_ncu.log 2>&1
ncu -i gpurun_out/r2b7_k.ncu-rep --page raw --csv > gpurun_out/r2b7_k.raw.csv 2>/dev/null
ncu -i gpurun_out/r2b7_k.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2b7_k.cuda.csv 2>/dev/null
rm -f gpurun_out/r2b7_k.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"record_finish" -s 1 -c 1 -o gpurun_out/r2b7_f -f \
    python tools/profile_frame.py cfg3 16777216 2 > gpurun_out/r2b7_ncu_f.log 2>&1
ncu -i gpurun_out/r2b7_f.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2b7_f.cuda.csv 2>/dev/null
rm -f gpurun_out/r2b7_f.ncu-rep
gzip -f gpurun_out/r2b7_k.cuda.csv gpurun_out/r2b7_f.cuda.csv
ls -la gpurun_out/r2b7*
