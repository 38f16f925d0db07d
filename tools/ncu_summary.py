#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, on the CPU box) into profiles/<tag>_ncu_summary.{json,md}.

    python tools/ncu_summary.py gpurun_out/prof_mlp.ncu-rep r01

Keeps, per profiled kernel: duration, DRAM bytes read/written (the `traffic` bench.py reports),
tensor-pipe activity, L2 throughput, registers, shared memory, executed instructions.
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "inst_executed",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "msecond": 1e-3,
         "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0, "Ghz": 1e9, "Mhz": 1e6, "cycle/nsecond": 1e9, "cycle/usecond": 1e6}


def main(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    out = []
    for r in rows[2:]:
        k = {"kernel": r[name_col].replace("<unnamed>::", "").replace("void ", "").split("(")[0]}
        for i, h in enumerate(hdr):
            if h in KEEP and r[i] != "":
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                k[KEEP[h]] = v * SCALE.get(units[i], 1.0)
        if "dram_read" in k and "dram_write" in k:
            k["dram_traffic_bytes"] = k["dram_read"] + k["dram_write"]
            if k.get("duration"):
                k["dram_gbs"] = k["dram_traffic_bytes"] / k["duration"] / 1e9
        out.append(k)
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
    json.dump({"report": os.path.basename(rep), "note": "ncu --set full --clock-control none; durations are under the profiler (cold cache, serialised): use shares, not absolutes", "kernels": out},
              open(os.path.join(root, f"{tag}_ncu_summary.json"), "w"), indent=1)
    with open(os.path.join(root, f"{tag}_ncu_summary.md"), "w") as f:
        f.write(f"# ncu --set full summary ({tag}, {os.path.basename(rep)})\n\n")
        f.write("| kernel | grid x block | regs | smem | duration us | DRAM rd GB | DRAM wr GB | DRAM GB/s | DRAM % | tensor pipe active % | L2 % | SM clk GHz |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for k in out:
            f.write("| {kernel} | {g:.0f} x {b:.0f} | {r:.0f} | {s:.0f} | {d:.1f} | {rd:.3f} | {wr:.3f} | {bw:.0f} | {dp:.1f} | {tp:.1f} | {l2:.1f} | {clk:.2f} |\n".format(
                kernel=k["kernel"], g=k.get("grid", 0), b=k.get("block", 0), r=k.get("registers", 0), s=k.get("smem_dynamic", 0),
                d=k.get("duration", 0) * 1e6, rd=k.get("dram_read", 0) / 1e9, wr=k.get("dram_write", 0) / 1e9, bw=k.get("dram_gbs", 0),
                dp=k.get("dram_pct", 0), tp=k.get("tensor_pipe_active_pct", 0), l2=k.get("l2_pct", 0), clk=k.get("sm_clock", 0) / 1e9))
    print(open(os.path.join(root, f"{tag}_ncu_summary.md")).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
