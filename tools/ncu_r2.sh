#!/bin/bash
# Round-2 evidence run (one GPU): plain runs first, then the ncu passes of the same commands.  The .ncu-rep files are far
# beyond what gpurun brings back (64 MiB), so the pages read here are exported as CSV on the box and the reports dropped.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/exchange_frame.py 8 24 > gpurun_out/r2_exchange_frame.log 2>&1 || exit 1
python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2_profile_frame.log 2>&1 || exit 1
python tools/profile_frame.py cfg3 16777216 2 > gpurun_out/r2_profile_frame_cfg3.log 2>&1 || exit 1
# every kernel of the second 2^25-object frame (the first frame allocates); launches per frame from the plain run
S5=$(( $(grep -o 'launches=[0-9]*' gpurun_out/r2_profile_frame.log | cut -d= -f2) / 2 ))
S3=$(( $(grep -o 'launches=[0-9]*' gpurun_out/r2_profile_frame_cfg3.log | cut -d= -f2) / 2 ))
ncu --set full --clock-control none -s $S5 -c $S5 -o gpurun_out/r2_cfg5_frame \
    python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2_ncu_frame.log 2>&1
# ... and of one 2^24-object log-normal frame with the ID-parity filter (config 3)
ncu --set full --clock-control none -s $S3 -c $S3 -o gpurun_out/r2_cfg3_frame \
    python tools/profile_frame.py cfg3 16777216 2 > gpurun_out/r2_ncu_frame_cfg3.log 2>&1
# the exchange kernels (records, pairs) with 8 destinations
ncu --set full --clock-control none -k regex:exchange_pass -s 2 -c 2 -o gpurun_out/r2_exchange \
    python tools/exchange_frame.py 8 24 > gpurun_out/r2_ncu_exchange.log 2>&1
# the dominant kernel (first record pass of the second frame) with its source page
ncu --set full --clock-control none --import-source on -k regex:radix_pass -s 8 -c 1 -o gpurun_out/r2_cfg5_sort_pass_src \
    python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2_ncu_pass_src.log 2>&1
for r in r2_cfg5_frame r2_cfg3_frame r2_exchange r2_cfg5_sort_pass_src; do
    ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
ncu -i gpurun_out/r2_cfg5_sort_pass_src.ncu-rep --page source --csv > gpurun_out/r2_cfg5_sort_pass_src.source.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/*.ncu-rep
gzip -f gpurun_out/r2_cfg5_sort_pass_src.source.csv
du -sh gpurun_out
