#!/bin/bash
# Round-2 (second half) evidence run on one GPU: plain runs first, then the ncu passes of the same commands.  The .ncu-rep
# files exceed what gpurun brings back, so the pages read here are exported as CSV / text on the box and the reports dropped.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2b_profile_frame.log 2>&1 || exit 1
python tools/profile_frame.py cfg3 16777216 2 > gpurun_out/r2b_profile_frame_cfg3.log 2>&1 || exit 1
S5=$(( $(grep -o 'launches=[0-9]*' gpurun_out/r2b_profile_frame.log | cut -d= -f2) / 2 ))
S3=$(( $(grep -o 'launches=[0-9]*' gpurun_out/r2b_profile_frame_cfg3.log | cut -d= -f2) / 2 ))
# every kernel of the second frame of each workload, with the source page of the kernels that matter
ncu --set full --clock-control none --import-source on -s $S5 -c $S5 -o gpurun_out/r2b_cfg5_frame -f \
    python tools/profile_frame.py cfg5 33554432 2 > gpurun_out/r2b_ncu_frame.log 2>&1
ncu --set full --clock-control none --import-source on -s $S3 -c $S3 -o gpurun_out/r2b_cfg3_frame -f \
    python tools/profile_frame.py cfg3 16777216 2 > gpurun_out/r2b_ncu_frame_cfg3.log 2>&1
for r in r2b_cfg5_frame r2b_cfg3_frame; do
    ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
# top source lines of the dominant kernels (launch index inside the captured frame)
idx() { python - "$1" "$2" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
names = [r[rows[0].index("Kernel Name")] for r in rows[2:]]
print(next((i for i, n in enumerate(names) if sys.argv[2] in n), -1))
PY
}
for k in encode_kernel radix_pass_kernel scan_groups_kernel pair_finish_kernel; do
    i=$(idx gpurun_out/r2b_cfg5_frame.raw.csv $k)
    [ "$i" -ge 0 ] && python tools/ncu_lines.py gpurun_out/r2b_cfg5_frame.ncu-rep $i 40 > gpurun_out/r2b_lines_cfg5_$k.txt 2>&1
done
for k in record_finish scan_runs_kernel scan_emit_kernel; do
    i=$(idx gpurun_out/r2b_cfg3_frame.raw.csv $k)
    [ "$i" -ge 0 ] && python tools/ncu_lines.py gpurun_out/r2b_cfg3_frame.ncu-rep $i 40 > gpurun_out/r2b_lines_cfg3_$k.txt 2>&1
done
rm -f gpurun_out/*.ncu-rep
# launch list of the bench command (per-launch times are cold-cache and serialised: the SHARE per kernel is what to compare)
python bench.py --steps 2 --warmup 1 --no-extra --cpu-budget 0.2 > gpurun_out/r2b_bench_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --cpu-budget 0.2 > gpurun_out/r2b_ncu_launches.log 2>&1
du -sh gpurun_out
