set -x
SAN_CASES="c3 c3cl c3r c3rs" tools/run_sanitizers.sh r2e
bash tools/c3_shape_sweep2.sh > gpurun_out/r2s_c3_shape_sweep2.txt 2>&1
cat gpurun_out/r2s_c3_shape_sweep2.txt
