# one `ncu --set full` launch each of the three selection kernels that follow k_refine_g in pmvs2's kernel time (config-3 scene)
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
# launch 48 of pre/post_process is the first (largest) expansion wave: 48 seed waves come first; k_set_ref_image runs once per filter round
for k in k_post_process k_set_ref_image k_pre_process; do
  skip=48; [ $k = k_set_ref_image ] && skip=0
  timeout 300 ncu --set full --clock-control none -k regex:$k -s $skip -c 1 -o gpurun_out/r2_${k}_full cmvs-pmvs_b200/bin/pmvs2 /tmp/pl_dtu48/ option.txt PSET > /dev/null 2> gpurun_out/r2_${k}_ncu.err; echo "$k rc=$?"
  ncu -i gpurun_out/r2_${k}_full.ncu-rep --page raw --csv > gpurun_out/r2_${k}_full_raw.csv 2>/dev/null
  python tools/ncu_summaries.py metrics gpurun_out/r2_${k}_full_raw.csv gpurun_out/r2_${k}_full_ncu_metrics.csv
done
