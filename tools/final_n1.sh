# final 1-GPU lines of the round (run under gpurun)
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py --steps 2 --warmup 1 > gpurun_out/b_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch2.log 2>&1
wc -c gpurun_out/final_bench.json gpurun_out/final_ref.json gpurun_out/launches_final.csv
