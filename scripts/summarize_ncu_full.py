#!/usr/bin/env python
"""Summarise `ncu --set full --import-source on` reports (gpurun_out/*.ncu-rep) into one JSON under profiles/: the headline
counters of the raw page plus, from the SASS source page, the share of warp samples each kernel spends in `mbarrier.try_wait`
spin branches (who waits for whom in a warp-specialised kernel) and the top stalled instructions.
Usage: python scripts/summarize_ncu_full.py <out-name> <label>::<report.ncu-rep> [...]      (no GPU needed)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    out_name = sys.argv[1]
    res = {"source": "ncu --set full --clock-control none --import-source on (B200); read with ncu -i ... --page raw / --page source "
                     "--print-source sass", "captures": []}
    for spec in sys.argv[2:]:
        label, path = spec.rsplit("::", 1)
        raw = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "raw", "--csv"]))))
        hdr = raw[0]
        rows = [dict(zip(hdr, r)) for r in raw[2:] if len(r) == len(hdr)]
        if not rows:
            continue
        last = rows[-1]                                   # the last profiled launch (warm)
        cap = {"case": label, "kernel": last.get("Kernel Name", "")[:70], "report": os.path.basename(path)}
        for k in WANT:
            if k in last:
                cap[k] = last[k]
        src = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "source", "--csv", "--print-source", "sass"]))))
        shdr, blocks, cur = None, [], None
        for r in src:
            if r and r[0] == "Kernel Name":
                cur = []
                blocks.append(cur)
            elif r and r[0] == "Address":
                shdr = r
            elif cur is not None and shdr and len(r) == len(shdr):
                cur.append(dict(zip(shdr, r)))
        if blocks:
            b = blocks[-1]
            tot = sum(int(x["# Samples"]) for x in b) or 1
            waits = sum(int(x["# Samples"]) for x in b if "BRA" in x["Source"] and int(x["stall_long_sb"]) > 0.8 * int(x["# Samples"]) > 12)
            membar = sum(int(x["stall_membar"]) for x in b)
            top = sorted(b, key=lambda x: -int(x["# Samples"]))[:8]
            stalls = [k for k in shdr if k.startswith("stall_") and "Not Issued" not in k]
            cap["warp_samples"] = tot
            cap["share_of_samples_in_mbarrier_wait_branches"] = round(waits / tot, 3)
            cap["share_of_samples_stalled_on_membar_fences"] = round(membar / tot, 3)
            cap["top_instructions"] = [{"sass": x["Source"].strip()[:60], "samples": int(x["# Samples"]), "executed": int(x["Instructions Executed"]),
                                        "main_stall": max(stalls, key=lambda k: int(x[k]))} for x in top]
        res["captures"].append(cap)
    path = os.path.join(ROOT, "profiles", out_name)
    json.dump(res, open(path, "w"), indent=1)
    for c in res["captures"]:
        print(c["case"], c.get("gpu__time_duration.sum"), "us  tensor", c.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
              " waits", c.get("share_of_samples_in_mbarrier_wait_branches"))


if __name__ == "__main__":
    main()
