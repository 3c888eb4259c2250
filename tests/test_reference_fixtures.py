"""The reference's own golden fixtures, reproduced byte for byte.

tests/data/** of the reference are Git-LFS pointers: the blobs are absent, but every pointer holds the SHA-256
and the size of its file.  The inputs come from a public algorithm (utils/src/gen_test_data.rs:98-168 driven by
tests/gen_test_scenes.py:13-31), the validation files from `gen_validation_data`
(utils/src/gen_test_data.rs:761-785, tests/gen_validation_data.py) run on the n = 10 000 input:

    0_layer_unsorted   = Default layer .extend(system_bounds, object_bounds)      <- tests/test_layer.rs:25-40
    1_layer_sorted     = .sort()                                                  <- :57-90
    2_layer_collisions = .scan()                                                  <- :92-124

Regenerating a file and hashing it is therefore an equality test against the bytes the Rust crate wrote:
  * CPU (`-m "not gpu"`): the ChaCha/PCG/gen_range restatement reproduces all seven inputs, and the ORACLE's
    extend / sort / scan of the n = 10 000 scene reproduce all three validation files -- this is what pins the
    oracle end to end with reference-produced data.
  * GPU (`-m gpu`): the CUDA path through the C ABI reproduces the same three files, and agrees with the oracle
    on the other six reference scenes.

Layouts that hash equal (found by enumeration: container version x "with / without object_bounds"): the two
layer files are SceneV1_1 with an EMPTY object_bounds vector (385 137 = 12 + 24 + 8 + 4 + 8 + 12 * 32 090 + 1),
the collisions file is SceneV1_2 with the 10 000 objects, the sorted layer, 65 866 pairs, no hits, no nearest.
The constants below are copied from the pointer files (data, not code); when /root/reference is mounted the
test also checks them against the pointers themselves.
"""
import hashlib
import importlib.util
import os
import re

import numpy as np
import pytest

from oracle import cpu_oracle as co

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location("bp_" + name, os.path.join(ROOT, "broadphase-rs_b200", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


rust_rand = _load("rust_rand")      # numpy-only modules of the package: loaded by path, the CUDA library is not needed
scene_io = _load("scene_io")

# tests/data/inputs/boxes-seed_0-d_1_1000-s_1_10-n_<n>.br_scene : (sha256, size)
INPUTS = {
    100: ("9b0c04de94bfb0f26bc7ef7b514a6a87f119cb45511dfdf17d4185a3a9124752", 2844),
    300: ("97f21bf06961f9d9c0e31a38e7c92f77d3d07130dabeee5237876b89e67371d5", 8444),
    1000: ("c42fffb943673b2b2b4691e87b40e5e2967ad59ce47ca7f5da73d26a6699a99a", 28044),
    3000: ("a3b90443b3dd4a207d9144ed34e02381f75fc544555afdf2ce7b069b4e661fbb", 84044),
    10000: ("08c4a789322c5cd8934c044cb53aaa8735bc7b5401a9fd10f22535f6cb706c92", 280044),
    30000: ("c70e9f16eb0208f233a215415e8f320a74b7c678dc6744b381652fb827222121", 840044),
    100000: ("738667c97d094782d4a58549272984795f5617b2484cb6f0ce9d5332f7064023", 2800044),
}
# tests/data/validation/<name>.br_scene : (sha256, size)
VALIDATION = {
    "0_layer_unsorted": ("1cc4cd962afce12917d60efe1dae82a0bc34d14ab36d40372061cd4554bc3402", 385137),
    "1_layer_sorted": ("f7291f585d3f15e9a11c8c491090792531cba329deb438aaaf5bacc16cc58894", 385137),
    "2_layer_collisions": ("4e31c66ded02dc6f7163b4ba02570a894810a9c93fb2ab04301ad42fcc146a7e", 1192082),
}
REF_DATA = "/root/reference/tests/data"


def _sha(b):
    return hashlib.sha256(b).hexdigest()


def _scene(n):
    return rust_rand.gen_boxes_reference(n, seed=0, density=1e-3, size_range=(1.0, 10.0))


def validation_files(sysb, bounds, ids, unsorted, sorted_, pairs):
    """The three files of `gen_validation_data` from (keys, ids) before / after the sort and the pair list."""
    none_b, none_i = np.zeros((0, 6), np.float32), np.zeros(0, np.uint32)
    f0 = scene_io.Scene(sysb, none_b, none_i, 0, unsorted[0], unsorted[1], False).to_bytes((1, 1))
    f1 = scene_io.Scene(sysb, none_b, none_i, 0, sorted_[0], sorted_[1], True).to_bytes((1, 1))
    f2 = scene_io.Scene(sysb, bounds, ids, 0, sorted_[0], sorted_[1], True, collisions=pairs).to_bytes((1, 2))
    return {"0_layer_unsorted": f0, "1_layer_sorted": f1, "2_layer_collisions": f2}


# ---- the constants are the pointers' (only where the reference is mounted) --------------------------------------

@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference checkout not mounted (GPU box)")
def test_constants_match_the_lfs_pointers():
    def pointer(path):
        txt = open(path).read()
        return re.search(r"sha256:([0-9a-f]{64})", txt).group(1), int(re.search(r"size (\d+)", txt).group(1))
    for n, want in INPUTS.items():
        assert pointer("%s/inputs/boxes-seed_0-d_1_1000-s_1_10-n_%06d.br_scene" % (REF_DATA, n)) == want
    for name, want in VALIDATION.items():
        assert pointer("%s/validation/%s.br_scene" % (REF_DATA, name)) == want


# ---- rand_core / rand_chacha / rand restatement ----------------------------------------------------------------

def test_chacha20_block_function_rfc7539_vector():
    # RFC 7539 section 2.3.2 uses a 32-bit counter + 96-bit nonce; with nonce word 0 = 0 the state layout coincides
    # with the 64-bit-counter variant for counter = 1 and the remaining nonce words in the stream-id slot -- here only
    # the all-zero-nonce case is reachable, so check the well-known all-zero key / counter 0 keystream head instead.
    words = rust_rand.chacha20_words(np.zeros(8, np.uint32), 0, 2)
    assert words[:4].tobytes().hex() == "76b8e0ada0f13d90405d6ae55386bd28"
    assert words[16:20].tobytes().hex() == "9f07e7be5551387a98ba977c732d080d"


@pytest.mark.parametrize("n", sorted(INPUTS))
def test_gen_boxes_reproduces_the_reference_input_fixture(n):
    sysb, bounds, ids = _scene(n)
    data = scene_io.Scene(sysb, bounds, ids).to_bytes((1, 0))
    assert len(data) == INPUTS[n][1]
    assert _sha(data) == INPUTS[n][0]


def test_gen_range_redraws_when_rounded_up_to_high():
    class Fixed:
        def __init__(self, w): self.w = list(w)
        def take(self, n): out, self.w = np.array(self.w[:n], np.uint32), self.w[n:]; return out
        def untake(self, w): self.w = list(w) + self.w
    # 0xffffffff >> 9 -> v = 1 - 2^-23; v * 1 + 2^24 rounds up to high = 2^24 + 1?  use low = 2^24 (spacing 2): res = high
    rng = Fixed([0xFFFFFFFF, 0x00000000, 0x80000000])
    out = rust_rand.gen_range_f32(rng, np.array([16777216.0, 0.0], np.float32), np.array([16777218.0, 2.0], np.float32))
    assert out.tolist() == [16777216.0, 1.0]      # first word rejected (res == high), second accepted, third -> 0.5 * 2


# ---- the oracle against the reference's validation files -------------------------------------------------------

@pytest.fixture(scope="module")
def oracle_run():
    sysb, bounds, ids = _scene(10000)
    layer = co.OracleLayer(co.INDEX64_3D, 4, 0)                 # `layer: Default::default()`: min_depth 0
    layer.extend(sysb, bounds, ids)
    assert not layer.sorted
    k0, i0 = layer.records()
    layer.sort()
    k1, i1 = layer.records()
    layer.scan()
    pairs = layer.collisions()
    return sysb, bounds, ids, (k0, i0.astype(np.uint32)), (k1, i1.astype(np.uint32)), pairs.astype(np.uint32)


@pytest.mark.parametrize("name", sorted(VALIDATION))
def test_oracle_reproduces_the_reference_validation_file(oracle_run, name):
    sysb, bounds, ids, unsorted, sorted_, pairs = oracle_run
    data = validation_files(sysb, bounds, ids, unsorted, sorted_, pairs)[name]
    assert len(data) == VALIDATION[name][1]
    assert _sha(data) == VALIDATION[name][0]


def test_oracle_par_paths_equal_the_validated_ones(oracle_run):
    # tests/test_layer.rs:74-90, 109-124: par_sort / par_scan must give the same files
    sysb, bounds, ids, unsorted, sorted_, pairs = oracle_run
    layer = co.OracleLayer(co.INDEX64_3D, 4, 0)
    layer.extend(sysb, bounds, ids)
    layer.par_sort()
    k, i = layer.records()
    assert (k == sorted_[0]).all() and (i == sorted_[1]).all()
    layer.par_scan()
    assert (layer.collisions() == pairs).all()


def test_pyref_reproduces_the_reference_validation_files(oracle_run):
    from oracle import pyref
    sysb, bounds, ids, unsorted, sorted_, pairs = oracle_run
    k, i = pyref.extend(2, 0, sysb, bounds, ids)
    assert (k == unsorted[0]).all() and (i == unsorted[1]).all()
    ks, is_ = pyref.sort_records(k, i)
    assert (ks == sorted_[0]).all() and (is_ == sorted_[1]).all()
    p, _raw = pyref.scan(2, ks, is_)
    assert (np.asarray(p, dtype=np.uint32).reshape(-1, 2) == pairs).all()


# ---- the CUDA path against the same files ----------------------------------------------------------------------

@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference_validation_files(bp):
    sysb, bounds, ids = _scene(10000)
    layer = bp.LayerBuilder().build(bp.Index64_3D, "u32")      # min_depth 0 like `Default`
    layer.extend(sysb, bounds, ids)
    assert not layer.sorted
    k0, i0 = layer.iter()
    layer.sort()
    assert layer.sorted
    k1, i1 = layer.iter()
    pairs = np.asarray(layer.scan()).copy()
    files = validation_files(sysb, bounds, ids, (k0.astype(np.uint64), i0.astype(np.uint32)),
                             (k1.astype(np.uint64), i1.astype(np.uint32)), pairs.astype(np.uint32))
    for name, (sha, size) in VALIDATION.items():
        assert len(files[name]) == size, name
        assert _sha(files[name]) == sha, name


@pytest.mark.gpu
@pytest.mark.parametrize("n", sorted(INPUTS))
def test_cuda_path_equals_oracle_on_every_reference_input(bp, n):
    sysb, bounds, ids = _scene(n)
    g = bp.LayerBuilder().build(bp.Index64_3D, "u32")
    o = co.OracleLayer(co.INDEX64_3D, 4, 0)
    g.extend(sysb, bounds, ids)
    o.extend(sysb, bounds, ids)
    gk, gi = g.iter()
    ok, oi = o.records()
    assert (gk.astype(np.uint64) == ok).all() and (gi.astype(np.uint64) == oi).all()
    g.sort()
    o.par_sort()
    gk, gi = g.iter()
    ok, oi = o.records()
    assert (gk.astype(np.uint64) == ok).all() and (gi.astype(np.uint64) == oi).all()
    gp = np.asarray(g.scan())
    o.par_scan()
    op = o.collisions()
    assert gp.shape == op.shape and (gp.astype(np.uint64) == op).all()
