# bench.py prints exactly one JSON line when pmvs2 is timed from a fresh process image (N = 1, and N = all visible GPUs under torchrun)
set -x
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_stage_x1.json 2> gpurun_out/r2_stage_x1.err; echo "rc=$?"; wc -l gpurun_out/r2_stage_x1.json
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_stage_x1.json").read().strip())
p=d["pipeline"]; print("N=1 value %.3f" % (d["value"]/1e6), "pipeline wall %.3f" % p["wall_seconds"], "main %.3f ctx %.3f" % (p["phases_seconds"]["main.total"], p["phases_seconds"]["load.create_gpu_context"]), p["patches"], p["measured_from"])
P
N=$(nvidia-smi -L | wc -l)
if [ "$N" -gt 1 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2_stage_xN.json 2> gpurun_out/r2_stage_xN.err; echo "rc=$?"; wc -l gpurun_out/r2_stage_xN.json
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_stage_xN.json").read().strip())
p=d["pipeline"]; print("N=%d value %.3f" % (d["n_gpus"], d["value"]/1e6), "pipeline wall %.3f" % p["wall_seconds"], "main %.3f ctx %.3f" % (p["phases_seconds"]["main.total"], p["phases_seconds"]["load.create_gpu_context"]), p["patches"], p.get("exchange"))
P
fi
