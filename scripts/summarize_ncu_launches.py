#!/usr/bin/env python
"""Turn the ncu launch-list CSV of `bench.py` (timed region selected with --nvtx --nvtx-include "sib_timed/") into the
committed summaries under profiles/: one row per launch, and per-kernel-family totals (time share, DRAM bytes, tensor-pipe
activity) that bench.py reads for `roofline.traffic`.
Usage: python scripts/summarize_ncu_launches.py gpurun_out/launches_vNN.csv vNN [steps] [round] [git_head] [workload] [source_digest]
The git head of the tree the capture was taken on is stamped into the summary (bench.py quotes it in roofline.traffic_note)."""
import collections
import csv
import json
import re
import sys

FAMILIES = ["conv1d_bf16_tc_kernel", "resunit_tc_kernel", "attention_tc_kernel", "layernorm", "conv0_gn_apply", "conv0_gn_stats",
            "conv1d_cout1", "cast", "znorm", "zero_ranges", "gather", "assign_kernel", "conv1d_f32_kernel", "paste",
            "extend_mel", "transpose", "pack_int16", "elementwise", "emset"]


def family(name):
    # template arguments: conv1d_bf16_tc_kernel<POST_ACT, PAIR, EPI, FLOW>, resunit_tc_kernel<PAIR, WIDE>
    m = re.search(r"(conv1d_bf16_tc_kernel|resunit_tc_kernel)<([^>]*)>", name)
    if m:
        args = [a.strip().split(")")[-1] for a in m.group(2).split(",")]
        pair = args[1] if m.group(1) == "conv1d_bf16_tc_kernel" and len(args) > 1 else args[0]
        return m.group(1), ("pair, cta_group::2" if pair == "1" else "single CTA")
    for k in FAMILIES:
        if k in name:
            return k, None
    return name[:40], None


def main():
    src, tag = sys.argv[1], sys.argv[2]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    rnd = sys.argv[4] if len(sys.argv) > 4 else "r02"
    head = sys.argv[5] if len(sys.argv) > 5 else "unknown"
    wkl = sys.argv[6] if len(sys.argv) > 6 else "cfg2"
    digest = sys.argv[7] if len(sys.argv) > 7 else "n/a"
    hdr, recs = None, collections.OrderedDict()
    for r in csv.reader(open(src)):
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            rec = recs.setdefault(d["ID"], {"kernel": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"]})
            v, u, m = float(d["Metric Value"].replace(",", "")), d["Metric Unit"], d["Metric Name"]
            if m == "gpu__time_duration.sum":
                rec["us"] = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
            elif m.startswith("dram"):
                rec[m] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            else:
                rec["tensor_pct"] = v
    agg = collections.OrderedDict()
    for r in recs.values():
        fam, var = family(r["kernel"])
        for key in ([fam] if var is None else [fam, f"{fam}[{var}]"]):
            a = agg.setdefault(key, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tw": 0.0})
            a["n"] += 1
            a["us"] += r.get("us", 0)
            a["rd"] += r.get("dram__bytes_read.sum", 0)
            a["wr"] += r.get("dram__bytes_write.sum", 0)
            a["tw"] += r.get("tensor_pct", 0) * r.get("us", 0)
    tot = sum(a["us"] for k, a in agg.items() if "[" not in k)
    out = []
    print(f"{'kernel':52s} {'launches':>8s} {'us/step':>9s} {'share':>6s} {'rd MB':>9s} {'wr MB':>9s} {'tensor %':>9s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        tp = a["tw"] / a["us"] if a["us"] else 0
        print(f"{k:52s} {a['n'] / steps:8.1f} {a['us'] / steps:9.1f} {100 * a['us'] / tot:5.1f}% {a['rd'] / steps / 1e6:9.1f} "
              f"{a['wr'] / steps / 1e6:9.1f} {tp:9.1f}")
        out.append({"kernel": k, "launches_per_step": a["n"] / steps, "us_per_step": round(a["us"] / steps, 1),
                    "share": round(a["us"] / tot, 4), "dram_read_bytes_per_step": a["rd"] / steps,
                    "dram_write_bytes_per_step": a["wr"] / steps, "tensor_pipe_active_pct_time_weighted": round(tp, 1)})
    json.dump({"source": "ncu --nvtx --nvtx-include sib_timed/ --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                         "dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none "
                         f"python bench.py --workload {wkl} --steps {steps} --warmup 3 --no-cpu-baseline (B200, {rnd} {tag})",
               "git_head": head, "source_digest": digest, "workload": wkl, "steps": steps, "kernels": out},
              open(f"profiles/{rnd}_ncu_step_{tag}_summary.json", "w"), indent=1)
    with open(f"profiles/{rnd}_ncu_launches_{tag}.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum[us]", "dram__bytes_read.sum", "dram__bytes_write.sum",
                    "sm__pipe_tensor_cycles_active.pct"])
        for k, r in recs.items():
            w.writerow([k, r["kernel"][:110], r["grid"], r["block"], round(r.get("us", 0), 2), int(r.get("dram__bytes_read.sum", 0)),
                        int(r.get("dram__bytes_write.sum", 0)), r.get("tensor_pct", "")])


if __name__ == "__main__":
    main()
