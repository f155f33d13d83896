# every kernel launch of one pmvs2 run on the config-3 scene (DTU-48): plain run first (phase clocks), then the same command
# under `ncu --metrics gpu__time_duration.sum --clock-control none` (launch list; times under ncu are never bench values)
set -x
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
cmvs-pmvs_b200/bin/pmvs2 $PFX option.txt PSET > /dev/null 2> gpurun_out/p_plain.err; echo "plain rc=$?"
grep "^time" gpurun_out/p_plain.err | sort -k3 -n -r | head -12
timeout 230 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/p_launches.csv cmvs-pmvs_b200/bin/pmvs2 $PFX option.txt PSET > /dev/null 2> gpurun_out/p_ncu.err; echo "ncu rc=$?"
wc -l gpurun_out/p_launches.csv
