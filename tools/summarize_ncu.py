#!/usr/bin/env python3
"""Summarise an ncu report (gpurun_out/*.ncu-rep) into profiles/: a JSON with the metrics the roofline discussion
uses and profiles/traffic.json (dram bytes per launch of the F=602 forward kernel, read by bench.py).

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/ncu_r1_full [--traffic]
"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__inst_executed.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct",
    "sm__inst_executed_pipe_lsu.sum",
]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return v


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    kernels = []
    for r in data:
        if len(r) < len(H):
            continue
        k = {"kernel": r[H.index("Kernel Name")][:120]}
        for m in WANT:
            if m in H:
                i = H.index(m)
                k[m] = {"value": to_float(r[i]), "unit": units[i]}
        kernels.append(k)
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    json.dump({"report": os.path.basename(rep), "kernels": kernels}, open(out + ".json", "w"), indent=1)
    if "--traffic" in sys.argv:
        def gb(k, m):
            v = k.get(m)
            if not v:
                return 0.0
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(v["unit"], 1)
            return v["value"] * scale
        # round 2: the dominant kernel is planned_gather_sum_kernel<5, ...> (F=602: five float4 chunks per lane);
        # one aggregation CALL is one launch per slab plus the row-padding copy, summed here per call
        wide = [k for k in kernels if "planned_gather_sum_kernel<5" in k["kernel"]]
        narrow = [k for k in kernels if "planned_gather_sum_kernel<1" in k["kernel"]]
        res = {}
        for tag, ks in (("fwd_F602", wide), ("F128", narrow)):
            if ks:
                t = sum(gb(k, "dram__bytes_read.sum") + gb(k, "dram__bytes_write.sum") for k in ks) / len(ks)
                res[tag] = {"bytes": t, "kernel": ks[0]["kernel"], "launches_averaged": len(ks),
                            "source": os.path.basename(rep) + " (ncu --set full, per launch; the bench's measured "
                                      "slab count is 1 on this graph, so one launch per call)"}
        if res:
            json.dump(res, open(os.path.join(os.path.dirname(out), "traffic.json"), "w"), indent=1)
    print("wrote", out + ".json", len(kernels), "kernels")


if __name__ == "__main__":
    main()
