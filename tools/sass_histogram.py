"""SASS instruction histogram per kernel of the built library (no GPU needed): python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess

SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "broadphase-rs_b200", "libbroadphase_b200.so")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur:
        op = m.group(1)
        hist[cur][op.split(".")[0]] += 1
        if op.startswith(("UBLKCP", "SYNCS", "UTMA", "ATOMS", "ATOMG", "RED", "VOTE", "MATCH", "LDL", "STL")):
            hist[cur]["~" + ".".join(op.split(".")[:3])] += 1
names = subprocess.run(["c++filt"], input="\n".join(hist.keys()), capture_output=True, text=True).stdout.splitlines()
print("SASS instruction histogram per kernel of libbroadphase_b200.so (sm_100a cubin; `cuobjdump -sass`, tools/sass_histogram.py).")
print("Columns: total instructions, then the mnemonics that matter for the claims in DESIGN.md: UBLKCP.S.G = cp.async.bulk global->shared (TMA load),")
print("UBLKCP.G.S = cp.async.bulk shared->global (TMA store), SYNCS = mbarrier ops, VOTE / MATCH = warp ranking, ATOMS = shared atomics, LDL / STL = spills.\n")
for (fn, h), dm in zip(hist.items(), names):
    tot = sum(v for k, v in h.items() if not k.startswith("~"))
    special = {k[1:]: v for k, v in h.items() if k.startswith("~")}
    name = re.sub(r"bp::", "", dm.split("(")[0].replace("void ", ""))[:110]
    print("%-112s %6d  %s" % (name, tot, " ".join("%s=%d" % (k, v) for k, v in sorted(special.items()))))
