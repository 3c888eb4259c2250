"""Top CUDA source lines of one launch by stall samples (needs -lineinfo and --import-source on).
Usage: python tools/ncu_lines.py rep [launch_idx] [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
li = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) == 2 and r[0] == "Function Name":
        print(r[1][:120])
    elif r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 6 and r[0].isdigit():
        try:
            lines.append((int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0), cur_file, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1
toti = sum(l[1] for l in lines) or 1
print("samples %d, warp instructions %d" % (tot, toti))
for smp, ins, f, ln, src in sorted(lines, key=lambda l: -l[0])[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * smp / tot, 100.0 * ins / toti, f, ln, src))
