"""BR_SCENE reader / writer (reference container: data/src/lib.rs:19-106)."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bp_scene_io", os.path.join(ROOT, "broadphase-rs_b200", "scene_io.py"))
scene_io = importlib.util.module_from_spec(spec)
spec.loader.exec_module(scene_io)

# `size` lines of the Git-LFS pointers of the reference's input fixtures, tests/data/inputs/
# boxes-seed_0-d_1_1000-s_1_10-n_<n>.br_scene (written by utils/src/gen_test_data.rs:98-168 as SceneV1_0 at the time)
REFERENCE_INPUT_SIZES = {100: 2844, 300: 8444, 1000: 28044, 3000: 84044, 10000: 280044, 30000: 840044, 100000: 2800044}


def _scene(n, seed=0):
    rng = np.random.Generator(np.random.Philox(seed))
    mn = rng.random((n, 3)).astype(np.float32) * 90
    b = np.concatenate([mn, mn + 1 + rng.random((n, 3)).astype(np.float32) * 9], axis=1).astype(np.float32)
    return scene_io.Scene([0, 0, 0, 100, 100, 100], b, np.arange(n, dtype=np.uint32))


@pytest.mark.parametrize("n,size", sorted(REFERENCE_INPUT_SIZES.items()))
def test_v1_0_files_have_the_sizes_of_the_reference_fixtures(n, size):
    assert len(_scene(n).to_bytes(version=(1, 0))) == size


@pytest.mark.parametrize("version", [(1, 0), (1, 1), (1, 2)])
def test_round_trip(version, tmp_path):
    sc = _scene(257, 3)
    sc.min_depth, sc.sorted = 3, True
    sc.keys = np.arange(1000, dtype=np.uint64) * np.uint64(0x0123456789AB)
    sc.rec_ids = (np.arange(1000) % 257).astype(np.uint32)
    sc.collisions = np.array([[5, 1], [9, 2]], dtype=np.uint32)
    sc.hits = np.array([4, 8, 15], dtype=np.uint32)
    sc.nearest = (16, 2.5)
    p = tmp_path / "s.br_scene"
    sc.save(p, version)
    back = scene_io.Scene.load(p)
    assert (back.system_bounds == sc.system_bounds).all() and (back.bounds == sc.bounds).all() and (back.ids == sc.ids).all()
    if version[1] >= 1:
        assert back.min_depth == 3 and back.sorted and (back.keys == sc.keys).all() and (back.rec_ids == sc.rec_ids).all()
    else:
        assert back.keys.shape == (0,) and not back.sorted          # From<SceneV1_0>: Default layer
    if version[1] >= 2:
        assert (back.collisions == sc.collisions).all() and (back.hits == sc.hits).all() and back.nearest == (16, 2.5)
    else:
        assert back.collisions.shape == (0, 2) and back.nearest is None


def test_rejects_bad_files():
    good = _scene(3).to_bytes()
    with pytest.raises(scene_io.SceneIOError):
        scene_io.Scene.from_bytes(b"NOTSCENE" + good[8:])
    with pytest.raises(scene_io.SceneIOError):
        scene_io.Scene.from_bytes(good[:8] + b"\x02\x00\x00\x00" + good[12:])     # major version 2
    with pytest.raises(scene_io.SceneIOError):
        scene_io.Scene.from_bytes(good[:8] + b"\x01\x00\x03\x00" + good[12:])     # minor version 3
    with pytest.raises(scene_io.SceneIOError):
        scene_io.Scene.from_bytes(good[:-3])
    with pytest.raises(scene_io.SceneIOError, match="LFS"):
        scene_io.Scene.from_bytes(b"version https://git-lfs.github.com/spec/v1\noid sha256:0\nsize 1\n")


def test_a_scene_file_drives_the_oracle_layer(tmp_path):
    """The use the reference makes of these files (tests/test_layer.rs:25-40): load, extend a layer, compare."""
    from oracle import cpu_oracle as co
    sc = _scene(500, 7)
    o = co.OracleLayer(co.INDEX64_3D, 4, 0)
    o.extend(sc.system_bounds, sc.bounds, sc.ids)
    sc.keys, sc.rec_ids = o.records()
    sc.sorted = bool(o.sorted)
    p = tmp_path / "validation.br_scene"
    sc.save(p)
    back = scene_io.Scene.load(p)
    o2 = co.OracleLayer(co.INDEX64_3D, 4, back.min_depth)
    o2.extend(back.system_bounds, back.bounds, back.ids)
    k2, i2 = o2.records()
    assert (k2 == back.keys).all() and (i2 == back.rec_ids).all() and bool(o2.sorted) == back.sorted
