"""Summarises `ncu --page raw --csv` exports (made on the GPU box by tools/ncu_r2.sh; the .ncu-rep files themselves exceed what
gpurun brings back): one line per launch, a per-kernel-class traffic table for profiles/traffic.json, top stall reasons.
Usage: python tools/ncu_csv_summary.py raw.csv [--json]"""
import csv
import json
import sys

CLASS = [("encode_kernel", "encode"), ("radix_hist_kernel", "hist"), ("radix_pass_kernel", "pass"), ("scan_runs_kernel", "scan_runs"),
         ("scan_emit_kernel", "scan_emit"), ("scan_groups_kernel", "scan_emit"), ("record_finish", "sort_finish"),
         ("pair_finish_kernel", "pair_unique"), ("pair_unique_kernel", "pair_unique"),
         ("merge_tiles_kernel", "merge"), ("exchange_pass_kernel", "partition"), ("partition_hist_kernel", "partition")]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, k):
        try:
            return float(r[idx[k]].replace(",", "")) * SCALE.get(units[idx[k]], 1.0)
        except Exception:
            return float("nan")
    return hdr, idx, data, val


def main():
    path = sys.argv[1]
    hdr, idx, data, val = load(path)
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    out = []
    seen_pass = 0
    print("%-34s %8s %9s %9s %6s %6s %6s %6s %5s %7s  %s" % ("kernel", "ms", "dram rd MB", "dram wr MB", "dram%", "issue%", "warps%", "bankcf%", "regs", "grid", "top stalls (warps per issue-active)"))
    for r in data:
        name = r[idx["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("bp::", "")[:34]
        ms = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        sh = val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        bc = val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
        st = sorted(stall, key=lambda h: -float(r[idx[h]] or 0))[:4]
        sts = " ".join("%s %.1f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[idx[h]] or 0)) for h in st)
        print("%-34s %8.3f %9.1f %9.1f %6.1f %6.1f %6.1f %6.1f %5d %7d  %s" % (
            short, ms, rd / 1e6, wr / 1e6, val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            100.0 * bc / sh if sh else 0.0, int(val(r, "launch__registers_per_thread")), int(val(r, "launch__grid_size")), sts))
        cls = next((c for k, c in CLASS if k in name), None)
        if cls == "hist" or cls == "pass":
            # the record sort comes first in a frame (one histogram, then its passes), the pair sort second
            if cls == "hist":
                seen_pass += 1
            cls = ("sort_" if seen_pass <= 1 else "pair_") + cls
        out.append((cls, rd + wr, ms))
    if "--json" in sys.argv:
        agg = {}
        for cls, b, ms in out:
            if cls:
                agg.setdefault(cls, []).append(b)
        print(json.dumps({c: int(sum(v) / len(v)) for c, v in agg.items()}, indent=1))


if __name__ == "__main__":
    main()
