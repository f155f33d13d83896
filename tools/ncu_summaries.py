"""Summaries of ncu output for profiles/ (the raw files stay in gpurun_out/, which is scratch).

  python tools/ncu_summaries.py launches LAUNCHES.csv "header text"        per-kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list
  python tools/ncu_summaries.py metrics RAW.csv OUT.csv [DRAM.json SRC]    selected metrics of the one kernel in an `ncu -i X.ncu-rep --page raw --csv` dump
                                                                            (+ the DRAM bytes per launch that bench.py quotes as roofline.traffic)
"""
import collections
import csv
import json
import re
import sys

PICK = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct", "smsp__issue_active.avg.pct", "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__f_tex2sm_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tex_wavefronts.sum", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_tex_mem_texture.sum",
    "l1tex__t_sectors_pipe_tex_mem_texture_lookup_hit.sum", "l1tex__t_sectors.sum", "l1tex__t_sectors_pipe_tex.sum", "l1tex__t_requests_pipe_tex.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def short_name(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("(anonymous namespace)", "<unnamed>")
    return name.strip()


def launches(path, header):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    total = collections.Counter(); count = collections.Counter()
    for r in rows[hdr + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        unit = r[mu]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v * 1e3 if unit in ("s", "second") else v
        k = short_name(r[kn])
        total[k] += ms; count[k] += 1
    tot = sum(total.values())
    print(header)
    print("launches %d, summed kernel time %.1f ms" % (sum(count.values()), tot))
    for k, ms in total.most_common():
        print("%9.2f ms %5.1f%% %5d  %s" % (ms, 100.0 * ms / tot, count[k], k))


def metrics(raw, out, dram_json=None, source=""):
    rows = [r for r in csv.reader(open(raw, errors="replace")) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    vals = rows[hdr + 2]
    got = {}
    for n, u, v in zip(names, units, vals):
        base = n.split(".", 2)[-1] if n.count(".") >= 2 and n.split(".")[0].isupper() else n      # "FBSP.TriageCompute.dram__..." -> metric name
        for p in PICK:
            if n == p or n.endswith("." + p) or base == p:
                got[p] = (u, v)
    with open(out, "w") as f:
        f.write("metric,unit,value\n")
        for p in sorted(got):
            f.write("%s,%s,%s\n" % (p, got[p][0], got[p][1].replace(",", "")))
    print("wrote", out, len(got), "metrics; kernel:", vals[names.index("Kernel Name")][:80])
    if dram_json:
        def tobytes(p):
            u, v = got[p]
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        rd, wr = tobytes("dram__bytes_read.sum"), tobytes("dram__bytes_write.sum")
        grid = int(float(got["launch__grid_size"][1]))
        json.dump({"kernel": "k_refine_g<7, atlas>", "patches_per_launch": 1048576, "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd,
                   "dram_bytes_write": wr, "grid": grid, "source": source}, open(dram_json, "w"), indent=1)
        print("wrote", dram_json, rd + wr)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        metrics(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None, sys.argv[5] if len(sys.argv) > 5 else "")
