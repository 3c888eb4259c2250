"""Generates the golden fixtures in this directory from the CPU oracle (oracle/bp_oracle.cpp).

The reference's own golden files are Git-LFS stubs and the Rust crate cannot be built in this image, so
these vectors are restatement-derived: they pin the oracle (and, through it, the CUDA path) against
regressions, they are NOT reference output.  Run from the repository root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "broadphase-rs_b200"))
import scenes  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "example_circles_2000": lambda: scenes.example_circles(2000, 1),
    "uniform_cubes_4096": lambda: scenes.uniform_cubes(4096, 2),
    "lognormal_cubes_4096": lambda: scenes.lognormal_cubes(4096, 3),
    "gen_boxes_1000": lambda: scenes.gen_boxes(1000, 0),
    "edge_cases_3d": scenes.edge_cases_3d,
}


def main():
    for name, make in CASES.items():
        sc = make()
        L = co.OracleLayer(sc["kind"], 4, sc["min_depth"])
        L.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
        uk, ui = L.records()
        L.sort()
        sk, si = L.records()
        pairs = L.scan()
        parity = L.scan(co.FILTER_ID_PARITY)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=sc["kind"], min_depth=sc["min_depth"],
                            sys_bounds=sc["sys_bounds"], bounds=sc["bounds"], ids=sc["ids"],
                            unsorted_keys=uk, unsorted_ids=ui.astype(np.uint32), sorted_keys=sk,
                            sorted_ids=si.astype(np.uint32), pairs=pairs.astype(np.uint32),
                            pairs_id_parity=parity.astype(np.uint32))
        print(name, "records", uk.shape[0], "pairs", pairs.shape[0], "parity", parity.shape[0])


if __name__ == "__main__":
    main()
