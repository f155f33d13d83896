# round-2 measurement: tests, bench line (both arms), ncu launch list of the bench command, one --set full capture of the hot kernel,
# the launch list of one pmvs2 run on the config-3 scene.  Numbers printed under ncu are never bench values.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_final_tests.log 2>&1; tail -4 gpurun_out/r2_final_tests.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/r2_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/r2_ncu_launch.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_refine -s 3 -c 1 -o gpurun_out/r2_refine_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/r2_ncu_full.log 2>&1; echo "full rc=$?"
ncu -i gpurun_out/r2_refine_full.ncu-rep --page raw --csv > gpurun_out/r2_refine_full_raw.csv 2>/dev/null
python - <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
synth = g.load_package().synth
scene = synth.dtu_scene()
synth.render(scene, device="cuda")
scene.option["CPU"] = os.cpu_count() or 4
print(synth.write_scene(scene, "/tmp/pl_dtu48"))
P
PFX=/tmp/pl_dtu48/
for i in 1 2 3; do cmvs-pmvs_b200/bin/pmvs2 $PFX option.txt PSET > /dev/null 2> gpurun_out/r2_pmvs2_dtu48_plain$i.err; echo "plain rc=$?"; grep "time main.total\|time load.create" gpurun_out/r2_pmvs2_dtu48_plain$i.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_pmvs2_dtu48_launches_ncu.csv cmvs-pmvs_b200/bin/pmvs2 $PFX option.txt PSET > /dev/null 2> gpurun_out/r2_pmvs2_ncu.err; echo "ncu rc=$?"
wc -l gpurun_out/r2_pmvs2_dtu48_launches_ncu.csv
