#!/bin/bash
# Where does pmvs2's start-up time go?  PMVSB_TRACE_INIT prints the CUDA start-up steps of pmvsb_create; run alone and
# while another process (python + torch) holds a context on the same GPU.  usage: tools/probe/init_probe.sh <scene-prefix>
P=$1
nvidia-smi --query-gpu=persistence_mode,name --format=csv,noheader
run() {
  local t0=$(date +%s.%N)
  PMVSB_TRACE_INIT=1 cmvs-pmvs_b200/bin/pmvs2 $P option.txt PSET > /dev/null 2> /tmp/init_probe.err
  local t1=$(date +%s.%N)
  echo "$1 wall $(echo "$t1 - $t0" | bc) s rc $?"
  grep -E "^init|time main.total|time load" /tmp/init_probe.err | tr '\n' ';'
  echo
}
for i in 1 2 3; do run alone-$i; done
python - <<PY &
import torch, time
x = torch.empty(1 << 28, device="cuda"); torch.cuda.synchronize()
time.sleep(30)
PY
sleep 12
for i in 1 2 3; do run with-parent-$i; done
wait
