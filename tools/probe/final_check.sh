# last pass of the round on one GPU: the whole GPU test suite, smoke(), the default bench line
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_final_tests.log 2>&1; tail -3 gpurun_out/r2_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; wc -l gpurun_out/r2_bench.json
