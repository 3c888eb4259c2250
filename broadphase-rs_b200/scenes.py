"""Synthetic scenes for BASELINE.json's configs, numpy only.

Every recipe is a pure function of (n, seed) built on numpy's counter-based Philox generator, so
the GPU path, the CPU oracle and the bench all see identical f32 bits.  The distributions follow
the reference's own generators where one exists:
  * example scene   -- examples/main.rs:163-169 (bounds), :314 (radius), :415-417 (AABB)
  * gen_boxes scene -- utils/src/gen_test_data.rs:98-168, tests/gen_test_scenes.py:13-15
"""
import numpy as np

INDEX32_2D, INDEX64_2D, INDEX64_3D = 0, 1, 2


def _rng(seed):
    return np.random.Generator(np.random.Philox(int(seed)))


def example_circles(n=10_000, seed=1):
    """Config 1: n circles in the example's [-1, 1280]^2 system, Index32_2D, min_depth 4."""
    rng = _rng(seed)
    sys_bounds = np.array([-1.0, -1.0, 1280.0, 1280.0], dtype=np.float32)
    r = np.exp(rng.uniform(0.5, 2.0, size=n)).astype(np.float32)
    cx = (r + rng.random(n, dtype=np.float32) * (np.float32(1280.0) - 2 * r)).astype(np.float32)
    cy = (r + rng.random(n, dtype=np.float32) * (np.float32(720.0) - 2 * r)).astype(np.float32)
    bounds = np.stack([cx - r, cy - r, cx + r, cy + r], axis=1).astype(np.float32)
    ids = np.arange(n, dtype=np.uint32)
    return dict(kind=INDEX32_2D, min_depth=4, sys_bounds=sys_bounds, bounds=bounds, ids=ids)


def uniform_cubes(n=1 << 20, seed=2, id_base=0, edge_factor=0.4):
    """Config 2 (and the static/dynamic layers of config 4, the shards of config 5):
    n cubes of edge 0.4 * n^(-1/3) uniformly placed in the unit cube, Index64_3D."""
    rng = _rng(seed)
    sys_bounds = np.array([0, 0, 0, 1, 1, 1], dtype=np.float32)
    s = np.float32(edge_factor * float(n) ** (-1.0 / 3.0))
    mn = (rng.random((n, 3), dtype=np.float32) * (np.float32(1.0) - s)).astype(np.float32)
    mx = np.minimum(mn + s, np.float32(1.0)).astype(np.float32)
    bounds = np.concatenate([mn, mx], axis=1)
    ids = (np.arange(n, dtype=np.uint64) + np.uint64(id_base)).astype(np.uint32)
    return dict(kind=INDEX64_3D, min_depth=0, sys_bounds=sys_bounds, bounds=bounds, ids=ids)


def lognormal_cubes(n=1 << 24, seed=3, sigma=0.6):
    """Config 3: cube edge 0.25 * n^(-1/3) * exp(sigma * N(0,1)) clamped to <= 0.25 (multi-depth
    keys), Index64_3D; meant for scan_filtered with the ID-parity filter."""
    rng = _rng(seed)
    sys_bounds = np.array([0, 0, 0, 1, 1, 1], dtype=np.float32)
    h = 0.25 * float(n) ** (-1.0 / 3.0)
    s = np.minimum(h * np.exp(sigma * rng.standard_normal(n, dtype=np.float32)), 0.25).astype(np.float32)
    mn = (rng.random((n, 3), dtype=np.float32) * (np.float32(1.0) - s[:, None])).astype(np.float32)
    mx = np.minimum(mn + s[:, None], np.float32(1.0)).astype(np.float32)
    bounds = np.concatenate([mn, mx], axis=1)
    ids = np.arange(n, dtype=np.uint32)
    return dict(kind=INDEX64_3D, min_depth=0, sys_bounds=sys_bounds, bounds=bounds, ids=ids)


def gen_boxes(n=10_000, seed=0, density=1e-3, size_range=(1.0, 10.0)):
    """The reference's test-scene recipe (utils/src/gen_test_data.rs:98-168): per-axis sizes in
    size_range, min uniform in [sys.min, sys.max - size], system edge cbrt(n/density) + avg size.
    (The reference seeds ChaCha; the stream here is Philox, so the boxes differ.)"""
    rng = _rng(seed)
    avg = np.float32((size_range[0] + size_range[1]) / 2.0)
    edge = np.float32(np.cbrt(np.float32(n) / np.float32(density)) + avg)
    sys_bounds = np.array([0, 0, 0, edge, edge, edge], dtype=np.float32)
    size = rng.uniform(size_range[0], size_range[1], size=(n, 3)).astype(np.float32)
    mn = (rng.random((n, 3), dtype=np.float32) * (edge - size)).astype(np.float32)
    mx = np.minimum(mn + size, edge).astype(np.float32)
    bounds = np.concatenate([mn, mx], axis=1)
    ids = np.arange(n, dtype=np.uint32)
    return dict(kind=INDEX64_3D, min_depth=0, sys_bounds=sys_bounds, bounds=bounds, ids=ids)


def gen_boxes_reference(n=10_000, seed=0, density=1e-3, size_range=(1.0, 10.0)):
    """The same recipe on the reference's OWN random stream (ChaChaRng::seed_from_u64 + gen_range, restated in
    rust_rand.py): for seed 0 / density 0.001 / sizes 1..10 these are, bit for bit, the scenes of the reference's
    tests/data/inputs/*.br_scene (SHA-256 checked in tests/test_reference_fixtures.py)."""
    try:
        from . import rust_rand
    except ImportError:      # loaded by path (tests, bench reference arm), not as a package member
        import importlib.util
        import os
        spec = importlib.util.spec_from_file_location("bp_rust_rand", os.path.join(os.path.dirname(os.path.abspath(__file__)), "rust_rand.py"))
        rust_rand = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(rust_rand)
    sys_bounds, bounds, ids = rust_rand.gen_boxes_reference(n, seed, density, size_range)
    return dict(kind=INDEX64_3D, min_depth=0, sys_bounds=sys_bounds, bounds=bounds, ids=ids)


def edge_cases_3d():
    """Hand-picked boxes in the [-64, 64]^3 system of the reference's own unit test
    (src/geom.rs:696-706) that exercise the quantiser's corner cases."""
    sys_bounds = np.array([-64, -64, -64, 64, 64, 64], dtype=np.float32)
    nan, inf = np.float32(np.nan), np.float32(np.inf)
    b = [
        [-32, -32, -32, 32, 32, 32],          # the reference's round-trip box: depth 1, 8 cells
        [-64, -64, -64, 64, 64, 64],          # the whole system: depth 0, key 0
        [0, 0, 0, 0, 0, 0],                   # zero extent at the centre (0x7fffff80)
        [64, 64, 64, 64, 64, 64],             # zero extent at the system maximum
        [-64, -64, -64, -64, -64, -64],       # zero extent at the system minimum
        [-1e-3, -1e-3, -1e-3, 1e-3, 1e-3, 1e-3],  # tiny box straddling the mid-line
        [63.99, 63.99, 63.99, 64, 64, 64],    # touching the maximum
        [-65, 0, 0, 1, 1, 1],                 # outside (min.x) -> rejected
        [0, 0, 0, 1, 1, 64.5],                # outside (max.z) -> rejected
        [nan, 0, 0, 1, 1, 1],                 # NaN passes contains(), quantises to 0
        [0, 0, 0, 1, nan, 1],                 # NaN max
        [-inf, 0, 0, 1, 1, 1],                # -inf -> rejected
        [1, 1, 1, 0.5, 0.5, 0.5],             # inverted box (undefined in the reference; wraps)
        [10, -20, 30, 10.5, -19.75, 30.125],  # anisotropic
        [-64, -64, -64, 0, 0, 0],             # exactly one octant (to the mid-line)
        [-64, -64, -64, -1e-5, -1e-5, -1e-5],
    ]
    bounds = np.array(b, dtype=np.float32)
    ids = np.arange(bounds.shape[0], dtype=np.uint32)
    return dict(kind=INDEX64_3D, min_depth=0, sys_bounds=sys_bounds, bounds=bounds, ids=ids)
