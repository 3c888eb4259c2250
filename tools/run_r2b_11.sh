set -x
python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2b11_profile_cfg5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"scan_groups|encode_kernel|radix_pass" -s 6 -c 6 -o gpurun_out/r2b11_k -f \
    python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2b11_ncu.log 2>&1
ncu -i gpurun_out/r2b11_k.ncu-rep --page raw --csv > gpurun_out/r2b11_k.raw.csv 2>/dev/null
for i in 0 1 5; do python tools/ncu_lines.py gpurun_out/r2b11_k.ncu-rep $i 45 > gpurun_out/r2b11_lines_$i.txt 2>&1; done
rm -f gpurun_out/r2b11_k.ncu-rep
