set -x
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2l_gputest.log 2>&1
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2l_smoke.log 2>&1
( time python bench.py ) > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
( time python bench.py --impl reference ) > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err
tail -3 gpurun_out/r2l_gputest.log gpurun_out/r2l_smoke.log gpurun_out/r2l_bench.err gpurun_out/r2l_bench_ref.err
