# the peer-memory exchange on one GPU: tests (ranks as processes sharing the device), kernel A/B against the previous build
set -x
if [ "$1" != "--ab-only" ]; then timeout 900 python -m pytest tests/test_gpu_peer.py -x -q --timeout 600 > gpurun_out/r2_p1_peer.log 2>&1; tail -3 gpurun_out/r2_p1_peer.log; fi
timeout 600 python tools/variant_bench.py prev.so libpmvs_b200.so prev.so libpmvs_b200.so 2>&1 | grep -v "^$" | tail -8
timeout 600 python tools/variant_bench.py --patches 1048576 --reps 2 prev.so libpmvs_b200.so 2>&1 | grep -v "^$" | tail -4
