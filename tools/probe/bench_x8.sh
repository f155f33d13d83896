# bench.py at N = 8 (peer-memory exchange of the records, pmvs2 with 8 ranks as the pipeline leg)
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_x8_peer.json 2> gpurun_out/r2_bench_x8_peer.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_x8_peer.json").read().strip().splitlines()[-1])
print("value %.3f M/s" % (d["value"]/1e6), "ms_per_step %.3f" % d["ms_per_step"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], d.get("exchange"))
p=d.get("pipeline",{})
print("pipeline wall", p.get("wall_seconds"), "patches", p.get("patches"), p.get("exchange"))
print({k:v for k,v in p.get("phases_seconds",{}).items() if k.startswith(("load","main","round","gpu.allgather","gpu.evaluate"))})
PY
tail -5 gpurun_out/r2_bench_x8_peer.err
