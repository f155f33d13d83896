"""Does another process's CUDA context on the same GPU slow pmvs2 down (bench.py runs it as a child while torch holds a
context)?  Runs pmvs2 on the DTU-48 scene three times: alone, alone again, and with this process holding a context and 1 GB.
usage: python tools/probe/pipeline_context_probe.py"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g


def run(prefix, tag):
    t = time.time()
    p = subprocess.run([os.path.join(ROOT, "cmvs-pmvs_b200/bin/pmvs2"), prefix, "option.txt", "PSET"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    ph = {l.split()[1]: float(l.split()[2]) for l in p.stderr.splitlines() if l.startswith("time ")}
    keys = ["main.total", "load.total", "load.create_gpu_context", "round.seed", "round.expand", "round.filter", "write.total", "gpu.append_table",
            "gpu.post_process", "gpu.refine", "gpu.upload_table+depth_maps", "host.sync_table", "evaluate.total"]
    print(tag, "wall %.2f" % (time.time() - t), " ".join("%s=%.3f" % (k.split(".", 1)[-1] if k.startswith("gpu.") else k, ph.get(k, -1)) for k in keys), flush=True)


def main():
    synth = g.load_package().synth
    scene = synth.dtu_scene()
    synth.render(scene, device="cpu")        # no CUDA context in this process yet
    scene.option["CPU"] = os.cpu_count() or 4
    prefix = synth.write_scene(scene, "/tmp/ctx_probe_dtu48")
    run(prefix, "alone-1")
    run(prefix, "alone-2")
    import torch
    x = torch.empty(1 << 28, dtype=torch.float32, device="cuda")   # 1 GB + a live context
    torch.cuda.synchronize()
    run(prefix, "parent-context-1")
    run(prefix, "parent-context-2")
    del x


if __name__ == "__main__":
    main()
