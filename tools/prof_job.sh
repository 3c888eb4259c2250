# Round-end measurement batch on one B200 (run under gpurun): bench lines, launch list, full ncu capture of the
# config-2 frame, the other configs, the query throughput.  Every ncu run follows a plain run of the same command.
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py --steps 2 --warmup 1 > gpurun_out/b_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch2.log 2>&1
python tools/profile_frame.py cfg2 > gpurun_out/plain_cfg2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"radix_pass|scan_runs|scan_emit|encode_kernel|pair_finish|pair_scatter|radix_hist|count_scan" -c 12 -o gpurun_out/prof_cfg2_final -f python tools/profile_frame.py cfg2 1048576 1 > gpurun_out/ncu_cfg2b.log 2>&1
python tools/profile_frame.py cfg3 > gpurun_out/plain_cfg3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"radix_pass" -c 2 -o gpurun_out/prof_cfg3_pass_final -f python tools/profile_frame.py cfg3 16777216 1 > gpurun_out/ncu_cfg3b.log 2>&1
python tools/bench_configs.py > gpurun_out/configs_final.json 2> gpurun_out/configs_final.err
python tools/bench_queries.py 20 20 > gpurun_out/queries_final.json 2> gpurun_out/queries_final.err
tail -2 gpurun_out/ncu_cfg3b.log
