# final 1-GPU lines of the round (run under gpurun)
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_final_gputests_n1.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> gpurun_out/r2_final_gputests_n1.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_final_bench_n1_reference_arm.json 2> gpurun_out/r2_final_bench_n1_reference_arm.err
cat gpurun_out/r2_final_gputests_n1.log
