set -x
tools/sort_bench 146000000 0x3fffffe00000000 kv stop > gpurun_out/r2b9_sortbench_27bit_rb7.log 2>&1
tools/sort_bench 87000000 0x3fffffc000000000 kv stop > gpurun_out/r2b9_sortbench_24bit.log 2>&1
cat gpurun_out/r2b9_sortbench_27bit_rb7.log gpurun_out/r2b9_sortbench_24bit.log
