"""BR_SCENE files -- the reference's scene container (data/src/lib.rs:19-106), read and written with numpy only.

SURVEY.md section 8f rank 3: lets the reference's own fixtures (tests/data/**, Git-LFS blobs that are not
materialised in this checkout) and any user scene be replayed through the device layer.  The container is
bincode 1.x with its default options -- little-endian, fixed-width integers, u64 sequence lengths, a 1-byte
Option tag -- over these serde structs:

    Header      { signature: [u8; 8] = b"BR_SCENE", version: (u16, u16) }                 data/src/lib.rs:19-26
    SceneV1_0   { system_bounds: Bounds<Point3<f32>>, object_bounds: Vec<(Bounds, u32)> }  :28-32
    SceneV1_1   { .., layer: Layer<Index64_3D, u32> }                                      :34-39
    SceneV1_2   { .., collisions: Vec<(u32, u32)>, hits: Vec<u32>, nearest: Option<(u32, f32)> }  :41-49
    Layer       { min_depth: u32, tree: (Vec<(Index64_3D, u32)>, bool) }  (the other fields are #[serde(skip)])
                                                                                           src/layer.rs:40-68
    Bounds      { min: Point3 { x, y, z }, max: Point3 }                                   src/geom.rs:83-87

Pin: a v1.0 file with n objects is 44 + 28 n bytes -- exactly the sizes the LFS pointers of the reference's
seven input fixtures record (tests/test_scene_io.py).  This module never touches the GPU: it is plain file I/O.
"""
import struct

import numpy as np

SIGNATURE = b"BR_SCENE"
VERSION = (1, 2)
_OBJ = np.dtype([("bounds", "<f4", (6,)), ("id", "<u4")])       # (Bounds<Point3<f32>>, u32): 28 bytes
_REC = np.dtype([("index", "<u8"), ("id", "<u4")])              # (Index64_3D, u32): 12 bytes
_PAIR = np.dtype([("a", "<u4"), ("b", "<u4")])


class SceneIOError(ValueError):
    """InvalidSignature / UnsupportedVersion / truncated file (data/src/lib.rs:53-59)."""


class Scene:
    """SceneV1_2 (data/src/lib.rs:41-51); older versions are upgraded with empty fields like `From<SceneV1_x>` (:108-133)."""

    def __init__(self, system_bounds, bounds, ids, min_depth=0, keys=None, rec_ids=None, sorted_=False, collisions=None,
                 hits=None, nearest=None):
        self.system_bounds = np.asarray(system_bounds, dtype=np.float32).reshape(6)
        self.bounds = np.asarray(bounds, dtype=np.float32).reshape(-1, 6)
        self.ids = np.asarray(ids, dtype=np.uint32).reshape(-1)
        self.min_depth = int(min_depth)
        self.keys = np.zeros(0, np.uint64) if keys is None else np.asarray(keys, dtype=np.uint64).reshape(-1)
        self.rec_ids = np.zeros(0, np.uint32) if rec_ids is None else np.asarray(rec_ids, dtype=np.uint32).reshape(-1)
        self.sorted = bool(sorted_)   # `#[derive(Default)]` leaves the flag false (src/layer.rs:40)
        self.collisions = np.zeros((0, 2), np.uint32) if collisions is None else np.asarray(collisions, dtype=np.uint32).reshape(-1, 2)
        self.hits = np.zeros(0, np.uint32) if hits is None else np.asarray(hits, dtype=np.uint32).reshape(-1)
        self.nearest = nearest        # None or (id, dist)

    # ---- Scene::assemble (data/src/lib.rs:91-100) ---------------------------------------------------------
    def to_bytes(self, version=VERSION):
        if version[0] != 1 or not 0 <= version[1] <= 2:
            raise SceneIOError("unsupported version %r" % (version,))
        out = [SIGNATURE, struct.pack("<HH", *version), self.system_bounds.astype("<f4").tobytes()]
        objs = np.zeros(self.ids.shape[0], dtype=_OBJ)
        objs["bounds"], objs["id"] = self.bounds, self.ids
        out += [struct.pack("<Q", objs.shape[0]), objs.tobytes()]
        if version[1] >= 1:
            recs = np.zeros(self.keys.shape[0], dtype=_REC)
            recs["index"], recs["id"] = self.keys, self.rec_ids
            out += [struct.pack("<I", self.min_depth), struct.pack("<Q", recs.shape[0]), recs.tobytes(),
                    struct.pack("<B", 1 if self.sorted else 0)]
        if version[1] >= 2:
            out += [struct.pack("<Q", self.collisions.shape[0]), self.collisions.astype("<u4").tobytes(),
                    struct.pack("<Q", self.hits.shape[0]), self.hits.astype("<u4").tobytes()]
            out.append(b"\x00" if self.nearest is None else b"\x01" + struct.pack("<If", int(self.nearest[0]), float(self.nearest[1])))
        return b"".join(out)

    def save(self, path, version=VERSION):
        with open(path, "wb") as f:
            f.write(self.to_bytes(version))

    # ---- Scene::parse (data/src/lib.rs:68-89) ----------------------------------------------------------------
    @staticmethod
    def from_bytes(data):
        pos = 0

        def take(n):
            nonlocal pos
            if pos + n > len(data):
                raise SceneIOError("truncated BR_SCENE file")
            b = data[pos:pos + n]
            pos += n
            return b

        sig = take(8)
        if sig != SIGNATURE:
            if sig.startswith(b"version "):
                raise SceneIOError("this is a Git-LFS pointer, not the scene itself")
            raise SceneIOError("invalid signature %r" % sig)
        version = struct.unpack("<HH", take(4))
        if version[0] != VERSION[0] or version[1] > VERSION[1]:
            raise SceneIOError("unsupported version %r" % (version,))
        sysb = np.frombuffer(take(24), dtype="<f4").copy()
        n = struct.unpack("<Q", take(8))[0]
        objs = np.frombuffer(take(n * _OBJ.itemsize), dtype=_OBJ)
        sc = Scene(sysb, objs["bounds"].copy(), objs["id"].copy())
        if version[1] >= 1:
            sc.min_depth = struct.unpack("<I", take(4))[0]
            r = struct.unpack("<Q", take(8))[0]
            recs = np.frombuffer(take(r * _REC.itemsize), dtype=_REC)
            sc.keys, sc.rec_ids = recs["index"].copy(), recs["id"].copy()
            sc.sorted = take(1) != b"\x00"
        if version[1] >= 2:
            p = struct.unpack("<Q", take(8))[0]
            sc.collisions = np.frombuffer(take(p * 8), dtype="<u4").reshape(-1, 2).copy()
            h = struct.unpack("<Q", take(8))[0]
            sc.hits = np.frombuffer(take(h * 4), dtype="<u4").copy()
            if take(1) != b"\x00":
                sc.nearest = struct.unpack("<If", take(8))
        return sc

    @staticmethod
    def load(path):
        with open(path, "rb") as f:
            return Scene.from_bytes(f.read())
