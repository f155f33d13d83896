# final state check on one box: GPU tests, smoke(), the default bench line (with pmvs2's phase clocks)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
