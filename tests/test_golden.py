"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle): the
oracle must keep reproducing them (CPU), and the CUDA path must reproduce them through the C ABI (GPU)."""
import glob
import os

import numpy as np
import pytest

from oracle import cpu_oracle as co

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def test_fixtures_exist():
    assert len(FIXTURES) >= 5


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    L = co.OracleLayer(int(g["kind"]), 4, int(g["min_depth"]))
    L.extend(g["sys_bounds"], g["bounds"], g["ids"])
    k, i = L.records()
    assert (k == g["unsorted_keys"]).all() and (i == g["unsorted_ids"]).all()
    L.par_sort()
    k, i = L.records()
    assert (k == g["sorted_keys"]).all() and (i == g["sorted_ids"]).all()
    assert (L.par_scan() == g["pairs"]).all()
    assert (L.scan(co.FILTER_ID_PARITY) == g["pairs_id_parity"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_cuda_path_reproduces_golden(bp, path):
    g = np.load(path)
    L = bp.LayerBuilder().with_min_depth(int(g["min_depth"])).build(int(g["kind"]), "u32")
    L.extend(g["sys_bounds"], g["bounds"], g["ids"])
    k, i = L.iter()
    assert (k.astype(np.uint64) == g["unsorted_keys"]).all() and (i == g["unsorted_ids"]).all()
    L.par_sort()
    k, i = L.iter()
    assert (k.astype(np.uint64) == g["sorted_keys"]).all() and (i == g["sorted_ids"]).all()
    p = L.par_scan()
    assert p.shape == g["pairs"].shape and (p == g["pairs"]).all()
    L.clear()
    L.extend(g["sys_bounds"], g["bounds"], g["ids"])   # fresh frame: the dedup-at-source path
    p = L.scan_filtered(bp.ScanFilter.id_parity())
    assert p.shape == g["pairs_id_parity"].shape and (p == g["pairs_id_parity"]).all()
