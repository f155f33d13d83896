# pmvs2 on the config-3 scene, three plain runs with the phase clocks (after the pipeline tests)
set -x
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -q --timeout 900 -k "cloud or two_ranks or variants or formats" > gpurun_out/r2_q_pipe.log 2>&1; tail -3 gpurun_out/r2_q_pipe.log
for i in 1 2 3; do
  timeout 300 python tools/compare_pipeline.py dtu48 --skip-ref --ranks 1 > gpurun_out/r2_q_dtu48_$i.json 2> gpurun_out/r2_q_dtu48_$i.err; cat gpurun_out/r2_q_dtu48_$i.json
  grep -h "^time" gpurun_out/pmvs2_dtu48.log | sed 's/^time //' | tr "\n" ";"; echo
done
