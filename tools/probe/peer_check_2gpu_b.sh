# second pass on two GPUs: the two forms of the record exchange against NCCL at N = 2 (bench), tests
set -x
timeout 900 python -m pytest tests/test_gpu_peer.py -q --timeout 600 > gpurun_out/r2_p3_peer.log 2>&1; tail -3 gpurun_out/r2_p3_peer.log
timeout 600 python tools/variant_bench.py --patches 1048576 --reps 2 libpmvs_b200.so 2>&1 | grep -v "^$" | tail -3
for mode in peer peer_in_kernel nccl; do
  export PMVSB_BENCH_GATHER=peer PMVSB_GATHER_IN_KERNEL=0
  [ $mode = nccl ] && export PMVSB_BENCH_GATHER=nccl
  [ $mode = peer_in_kernel ] && export PMVSB_GATHER_IN_KERNEL=1
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-pipeline > gpurun_out/r2_p3_bench_x2_$mode.json 2> gpurun_out/r2_p3_bench_x2_$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_p3_bench_x2_$mode.json").read().strip().splitlines()[-1])
print("$mode", "value %.4f M/s" % (d["value"]/1e6), "ms_per_step %.3f" % d["ms_per_step"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], d.get("exchange",{}).get("mode"), d.get("exchange",{}).get("verified_against_nccl_allgather"), "launches", d["gpu_launches"])
PY
done
