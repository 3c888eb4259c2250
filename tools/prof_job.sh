set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench18.json 2> gpurun_out/bench18.err
python bench.py --steps 2 --warmup 1 > gpurun_out/b_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch2.log 2>&1
python tools/profile_frame.py cfg2 > gpurun_out/plain_cfg2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"radix_pass|scan_runs|scan_emit|encode_kernel|pair_finish|pair_scatter|radix_hist" -c 12 -o gpurun_out/prof_cfg2_r1b -f python tools/profile_frame.py cfg2 1048576 1 > gpurun_out/ncu_cfg2b.log 2>&1
tail -2 gpurun_out/ncu_cfg2b.log
