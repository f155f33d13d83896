# the peer-memory exchange between two GPUs: tests, pmvs2 on DTU-48 / Ring-47 with 1 and 2 ranks (peer and nccl), bench at N = 2
set -x
nvidia-smi topo -m 2>&1 | head -8
timeout 900 python -m pytest tests/test_gpu_peer.py -q --timeout 600 > gpurun_out/r2_p2_peer.log 2>&1; tail -5 gpurun_out/r2_p2_peer.log
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -q --timeout 900 -k "two_gpus or cloud" > gpurun_out/r2_p2_pipe.log 2>&1; tail -5 gpurun_out/r2_p2_pipe.log
for s in dtu48 ring47; do
  timeout 300 python tools/compare_pipeline.py $s --skip-ref --ranks 1 > gpurun_out/r2_p2_${s}_x1.json 2> gpurun_out/r2_p2_${s}_x1.err; cat gpurun_out/r2_p2_${s}_x1.json; grep -h "^time\|^exchange" gpurun_out/pmvs2_${s}.log | tr "\n" ";"; echo
  cp gpurun_out/pmvs2_${s}.log gpurun_out/r2_p2_pmvs2_${s}_x1.log
  timeout 300 python tools/compare_pipeline.py $s --skip-ref --ranks 2 > gpurun_out/r2_p2_${s}_x2.json 2> gpurun_out/r2_p2_${s}_x2.err; cat gpurun_out/r2_p2_${s}_x2.json; grep -h "^time\|^exchange" gpurun_out/pmvs2_${s}_x2.log | tr "\n" ";"; echo
  cp gpurun_out/pmvs2_${s}_x2.log gpurun_out/r2_p2_pmvs2_${s}_x2_peer.log
  PMVSB_EXCHANGE=nccl timeout 300 python tools/compare_pipeline.py $s --skip-ref --ranks 2 > gpurun_out/r2_p2_${s}_x2n.json 2> gpurun_out/r2_p2_${s}_x2n.err; cat gpurun_out/r2_p2_${s}_x2n.json; grep -h "^time\|^exchange" gpurun_out/pmvs2_${s}_x2.log | tr "\n" ";"; echo
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-pipeline > gpurun_out/r2_p2_bench_x2.json 2> gpurun_out/r2_p2_bench_x2.err; cat gpurun_out/r2_p2_bench_x2.json | cut -c1-1500; tail -3 gpurun_out/r2_p2_bench_x2.err
PMVSB_BENCH_GATHER=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 5 --warmup 3 --no-pipeline > gpurun_out/r2_p2_bench_x2n.json 2> gpurun_out/r2_p2_bench_x2n.err; cat gpurun_out/r2_p2_bench_x2n.json | cut -c1-400
