"""Summarises an .ncu-rep (read here, without a GPU): headline metrics per launch and stall samples
per barrier-delimited code region of one launch.  Usage: python tools/ncu_summary.py rep [launch_idx]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
li = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "smsp__inst_executed.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
for w in want:
    if w in idx:
        print("%-70s %-14s %s" % (w, units[idx[w]], [r[idx[w]][:60] for r in data]))
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
print("stalls (warps per issue-active) for launch", li)
for h in sorted(stall, key=lambda h: -float(data[li][idx[h]] or 0))[:8]:
    print("   %-40s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), data[li][idx[h]]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
seen, uniq = set(), []
for r in data:
    if r[0] not in seen:
        seen.add(r[0])
        uniq.append(r)
data = uniq


def g(r, k):
    try:
        return int(float(r[idx[k]] or 0))
    except Exception:
        return 0


tot = sum(g(r, "# Samples") for r in data)
print("source page: %d instructions, %d samples" % (len(data), tot))
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
segs, start = [], 0
for n, r in enumerate(data):
    if "BAR.SYNC" in r[idx["Source"]] or "EXIT" in r[idx["Source"]]:
        segs.append((start, n))
        start = n + 1
segs.append((start, len(data) - 1))
for a, b in segs:
    s = sum(g(r, "# Samples") for r in data[a:b + 1])
    if s < tot * 0.01:
        continue
    br = {k.replace("stall_", ""): sum(g(r, k) for r in data[a:b + 1]) for k in keys}
    br = {k: v for k, v in br.items() if v > s * 0.08}
    ops = {}
    for r in data[a:b + 1]:
        t = r[idx["Source"]].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:7]
    print("  [%4d,%4d] %5.1f%% %s %s" % (a, b, 100.0 * s / tot, br, top))
hot = sorted(data, key=lambda r: -g(r, "# Samples"))[:12]
print("hottest instructions:")
for r in hot:
    print("  %5.1f%%  %s" % (100.0 * g(r, "# Samples") / tot, r[idx["Source"]].strip()[:80]))
