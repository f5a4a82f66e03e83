set -x
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r3s_gputest.log 2>&1
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r3s_smoke.log 2>&1
( time python bench.py ) > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err
( time python bench.py --impl reference ) > gpurun_out/r3s_bench_ref.json 2> gpurun_out/r3s_bench_ref.err
for f in gpurun_out/r3s_gputest.log gpurun_out/r3s_smoke.log gpurun_out/r3s_bench.err gpurun_out/r3s_bench_ref.err; do tail -n 5 $f; done
