"""Registers, stack (spill frame), shared and constant memory of every kernel of the built library, from
`cuobjdump --dump-resource-usage` (no GPU needed); names demangled and cut to the kernel + its leading template arguments.
   python tools/resource_usage.py > profiles/<round>_resource_usage.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "broadphase-rs_b200", "libbroadphase_b200.so")


def main():
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", SO], capture_output=True, text=True).stdout
    rows = []
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function\s+(\S+):", line)
        if m:
            name = m.group(1)
            continue
        if name and "REG:" in line:
            f = dict(kv.split(":") for kv in line.split() if ":" in kv)
            rows.append((name, int(f.get("REG", 0)), int(f.get("STACK", 0)), int(f.get("SHARED", 0)), int(f.get("LOCAL", 0))))
            name = None
    dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print("%-110s %5s %6s %7s %6s" % ("kernel (demangled, truncated)", "REG", "STACK", "SHARED", "LOCAL"))
    table = []
    for (mangled, reg, stack, shared, local), d in zip(rows, dem):
        d = re.sub(r"^void ", "", d)
        d = re.sub(r"\(.*$", "", d)          # drop the parameter list
        d = d.replace("bp::", "").replace("(anonymous namespace)::", "")
        table.append((d[:110], reg, stack, shared, local))
    for t in sorted(table):
        print("%-110s %5d %6d %7d %6d" % t)
    spilling = [t for t in table if t[2] > 0]
    print("\n%d kernels; %d with a stack frame (spills or local arrays): %s" %
          (len(table), len(spilling), ", ".join(sorted({t[0].split("<")[0] for t in spilling})) or "none"))


if __name__ == "__main__":
    sys.exit(main())
