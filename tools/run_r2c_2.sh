# round 2, third session: the whole 1-GPU test tier + smoke on the rebuilt library
set -x
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2c_gputests_n1.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> gpurun_out/r2c_gputests_n1.log
cat gpurun_out/r2c_gputests_n1.log
