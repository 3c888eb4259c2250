#!/bin/bash
# Round-2 evidence run (one GPU): plain runs first, then the ncu passes of the same commands.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --no-extra --steps 5 --warmup 3 --cpu-budget 1 > gpurun_out/r2_bench_6.json 2> gpurun_out/r2_bench_6.err || exit 1
python tools/exchange_frame.py 8 24 > gpurun_out/r2_exchange_frame.log 2>&1 || exit 1
python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2_profile_frame.log 2>&1 || exit 1
# launch list of the bench command (per-launch times are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --no-extra --steps 5 --warmup 3 --cpu-budget 1 > gpurun_out/r2_ncu_bench.log 2>&1
# every kernel of one 2^25-object frame
ncu --set full --clock-control none -s 30 -c 30 -o gpurun_out/r2_cfg5_frame \
    python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2_ncu_frame.log 2>&1
# the exchange kernels (records, pairs) with 8 destinations
ncu --set full --clock-control none -k regex:exchange_pass -s 2 -c 2 -o gpurun_out/r2_exchange \
    python tools/exchange_frame.py 8 24 > gpurun_out/r2_ncu_exchange.log 2>&1
# the dominant kernel with its source page
ncu --set full --clock-control none --import-source on -k regex:radix_pass -s 4 -c 1 -o gpurun_out/r2_cfg5_sort_pass_src \
    python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2_ncu_pass_src.log 2>&1
# the reports are far beyond what gpurun brings back (64 MiB): export what is read here and drop them
for r in r2_cfg5_frame r2_exchange r2_cfg5_sort_pass_src; do
    ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
ncu -i gpurun_out/r2_cfg5_sort_pass_src.ncu-rep --page source --csv > gpurun_out/r2_cfg5_sort_pass_src.source.csv 2>/dev/null
ncu -i gpurun_out/r2_exchange.ncu-rep --page details --csv > gpurun_out/r2_exchange.details.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/*.ncu-rep
gzip -f gpurun_out/r2_cfg5_sort_pass_src.source.csv
du -sh gpurun_out
