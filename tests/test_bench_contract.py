"""bench.py's output contract: exactly one JSON line on the process's ORIGINAL stdout, library chatter on stderr --
also when bench.py runs as __main__ under torchrun and dist_bench imports it a second time as module `bench`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import runpy
# what bench.main() does first, in the __main__ copy of the module
sys.stdout.flush()
real = os.dup(1)
os.environ["BP_BENCH_STDOUT_FD"] = str(real)
os.dup2(2, 1)
import bench                      # the second copy, as dist_bench.run imports it
print("NCCL version banner and other chatter")
bench.emit_line({"metric": "objects/sec for extend+sort+scan", "value": 1.0, "n_gpus": 2})
""" % ROOT


def test_json_line_reaches_the_original_stdout_through_a_second_import():
    r = subprocess.run([sys.executable, "-c", WORKER], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2
    assert "chatter" in r.stderr and "chatter" not in r.stdout
