# round-end measurement: tests, bench line, ncu launch list of the bench command, one full-size --set full capture
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/m_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/m_ncu_launch.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_refine -s 3 -c 1 -o gpurun_out/m_refine_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/m_ncu_full.log 2>&1; echo "full rc=$?"
