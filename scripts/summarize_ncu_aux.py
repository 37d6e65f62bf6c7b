#!/usr/bin/env python
"""Summarise the ncu CSV of scripts/aux_microbench.py (bandwidth-bound kernels) into profiles/<round>_ncu_aux_<tag>_summary.json:
per kernel the launch duration, DRAM bytes, DRAM throughput as a fraction of the ncu peak AND of the measured copy peak
(MEASURED_PEAKS.json hbm_gbs), plus the counters that say what binds the kernel when it is not HBM (issue slots, FMA / XU
pipes, LSU wavefronts).  Usage: python scripts/summarize_ncu_aux.py gpurun_out/aux_ncu_vN.csv vN [round] [git_head]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}


def main():
    src, tag = sys.argv[1], sys.argv[2]
    rnd = sys.argv[3] if len(sys.argv) > 3 else "r02"
    head = sys.argv[4] if len(sys.argv) > 4 else "unknown"
    hdr, recs = None, collections.OrderedDict()
    for r in csv.reader(open(src)):
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            rec = recs.setdefault(d["ID"], {"kernel": re.sub(r"\(.*", "", d["Kernel Name"])[:80], "grid": d["Grid Size"], "block": d["Block Size"]})
            v = float(d["Metric Value"].replace(",", "") or 0)
            rec[d["Metric Name"]] = v * UNIT.get(d["Metric Unit"], 1)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    fam = collections.OrderedDict()
    for r in recs.values():
        fam.setdefault((r["kernel"], r["grid"]), []).append(r)
    out = []
    for (k, grid), rs in fam.items():
        r = min(rs, key=lambda x: x.get("gpu__time_duration.sum", 1e30))   # best launch of this shape (first ones are cold)
        us = r.get("gpu__time_duration.sum", 0.0)
        byt = r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
        gbs = byt / us / 1e3 if us else 0.0
        out.append({"kernel": k, "grid": grid, "block": r["block"], "launches_seen": len(rs), "us": round(us, 2),
                    "dram_bytes": int(byt), "dram_gbs": round(gbs, 1), "dram_frac_of_measured_copy_peak": round(gbs / peaks["hbm_gbs"], 3),
                    "dram_throughput_pct_of_ncu_peak": round(r.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", 0.0), 1),
                    "issue_active_pct": round(r.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0), 1),
                    "fma_pipe_pct": round(r.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 0.0), 1),
                    "xu_pipe_pct": round(r.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 0.0), 1),
                    "lsu_wavefronts_pct": round(r.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 0.0), 1),
                    "warps_active_pct": round(r.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0), 1),
                    "sm_throughput_pct": round(r.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0), 1)})
    path = os.path.join(ROOT, "profiles", f"{rnd}_ncu_aux_{tag}_summary.json")
    json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum,dram__throughput...,smsp__issue_active...,pipe_fma,pipe_xu,"
                         "lsu wavefronts,warps_active,sm__throughput --clock-control none python scripts/aux_microbench.py (B200)",
               "git_head": head, "hbm_gbs_measured_copy_peak": peaks["hbm_gbs"], "kernels": out}, open(path, "w"), indent=1)
    print(f"{'kernel':46s} {'grid':>14s} {'us':>8s} {'GB/s':>8s} {'of copy':>8s} {'issue%':>7s} {'fma%':>6s} {'xu%':>6s} {'lsu%':>6s}")
    for o in out:
        print(f"{o['kernel'][:46]:46s} {o['grid']:>14s} {o['us']:8.1f} {o['dram_gbs']:8.0f} {o['dram_frac_of_measured_copy_peak']:8.2f} "
              f"{o['issue_active_pct']:7.1f} {o['fma_pipe_pct']:6.1f} {o['xu_pipe_pct']:6.1f} {o['lsu_wavefronts_pct']:6.1f}")


if __name__ == "__main__":
    main()
