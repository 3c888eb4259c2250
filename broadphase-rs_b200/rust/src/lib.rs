//! `Layer<Index, ID>` of zvxryb/broadphase-rs (src/layer.rs:42-68) over libbroadphase_b200.so.
//!
//! SOURCE ONLY -- never compiled in the build image (no cargo/rustc).  Every `extern "C"` item below is
//! declared in include/bp.h; the method names, argument meaning and implicit behaviour follow the crate.

use cgmath::{Point2, Point3};
use std::marker::PhantomData;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct BpLayerConfig {
    pub index_kind: i32,
    pub id_bytes: i32,
    pub min_depth: u32,
    pub device: i32,
    pub index_capacity: usize,
    pub collision_capacity: usize,
    pub test_capacity: usize,
}

#[repr(C)]
pub struct BpFilter {
    pub kind: i32,
    pub table_on_device: i32,
    pub arg: u64,
    pub table: *const u32,
    pub n_table: usize,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct BpPickResult {
    pub dist: f32,
    pub hit: u32,
    pub id: u64,
    pub point: [f32; 3],
    pub pad: u32,
}

/// Device functors standing in for pick_ray's `get_dist` closure (src/layer.rs:431-436): a table of shapes indexed by ID.
pub enum PickShapes<'a> {
    /// rows of `DIM + 1` floats: centre.., radius (the closure of examples/main.rs:427-449)
    Spheres(&'a [f32]),
    /// rows of `2 * DIM` floats: min.., max..
    Aabbs(&'a [f32]),
}

#[repr(C)]
pub struct BpLayer {
    _private: [u8; 0],
}

extern "C" {
    fn bp_layer_create(cfg: *const BpLayerConfig, out: *mut *mut BpLayer) -> c_int;
    fn bp_layer_destroy(layer: *mut BpLayer) -> c_int;
    fn bp_layer_clear(layer: *mut BpLayer) -> c_int;
    fn bp_layer_extend_host(layer: *mut BpLayer, system_bounds: *const f32, bounds: *const f32, ids: *const c_void, n: usize) -> c_int;
    fn bp_layer_merge(layer: *mut BpLayer, other: *const BpLayer) -> c_int;
    fn bp_layer_sort(layer: *mut BpLayer) -> c_int;
    fn bp_layer_scan(layer: *mut BpLayer, filter: *const BpFilter, out_pairs: *mut *const c_void, out_count: *mut usize) -> c_int;
    fn bp_layer_records(layer: *mut BpLayer, keys: *mut *const c_void, ids: *mut *const c_void, n: *mut usize, sorted: *mut c_int) -> c_int;
    fn bp_layer_set_records(layer: *mut BpLayer, keys: *const c_void, ids: *const c_void, n: usize, sorted: c_int, on_device: c_int) -> c_int;
    fn bp_layer_min_depth(layer: *const BpLayer, out_min_depth: *mut u32) -> c_int;
    fn bp_layer_test_box_batch(layer: *mut BpLayer, system_bounds: *const f32, boxes: *const f32, n_queries: usize, max_depth: i32,
                               on_device: c_int, out_pairs: *mut *const c_void, out_offsets: *mut *const u32, out_count: *mut usize) -> c_int;
    fn bp_layer_test_ray_batch(layer: *mut BpLayer, system_bounds: *const f32, rays: *const f32, n_queries: usize, max_depth: i32,
                               on_device: c_int, out_pairs: *mut *const c_void, out_offsets: *mut *const u32, out_count: *mut usize) -> c_int;
    fn bp_layer_pick_ray_batch(layer: *mut BpLayer, system_bounds: *const f32, rays: *const f32, n_queries: usize, max_dist: f32,
                               max_depth: i32, shape_kind: i32, shapes: *const f32, n_shapes: usize, on_device: c_int,
                               out_results: *mut *const BpPickResult) -> c_int;
    fn bp_layer_last_error(layer: *const BpLayer) -> *const c_char;
}

/// src/index.rs:293-295
pub trait SpatialIndex: Copy {
    const KIND: i32;
    const DIM: usize;
    type Key: Copy + PartialEq;
    type Point;
}
#[derive(Clone, Copy, Debug, Default, Eq, Ord, PartialEq, PartialOrd)]
pub struct Index32_2D(pub u32);
#[derive(Clone, Copy, Debug, Default, Eq, Ord, PartialEq, PartialOrd)]
pub struct Index64_2D(pub u64);
#[derive(Clone, Copy, Debug, Default, Eq, Ord, PartialEq, PartialOrd)]
pub struct Index64_3D(pub u64);
impl SpatialIndex for Index32_2D { const KIND: i32 = 0; const DIM: usize = 2; type Key = u32; type Point = Point2<f32>; }
impl SpatialIndex for Index64_2D { const KIND: i32 = 1; const DIM: usize = 2; type Key = u64; type Point = Point2<f32>; }
impl SpatialIndex for Index64_3D { const KIND: i32 = 2; const DIM: usize = 3; type Key = u64; type Point = Point3<f32>; }

/// src/traits.rs:6-16 -- the device path supports 32- and 64-bit integer IDs.
pub trait ObjectID: Copy + Ord + std::hash::Hash + std::fmt::Debug {
    const BYTES: i32;
}
impl ObjectID for u32 { const BYTES: i32 = 4; }
impl ObjectID for u64 { const BYTES: i32 = 8; }

/// src/geom.rs:84-87; `#[repr(C)]` so that a slice of bounds is `n x 2*DIM` floats (min.., max..).
#[repr(C)]
#[derive(Copy, Clone, Debug, PartialEq)]
pub struct Bounds<Point> {
    pub min: Point,
    pub max: Point,
}

/// Device functors standing in for `F: FnMut(ID, ID) -> bool` (src/layer.rs:456-460).
pub enum Filter<'a> {
    None,
    IdParity,
    XorMask(u64),
    Category(&'a [[u32; 2]]),
    /// fused narrow phase: `[x, y, z, r]` per ID; only pairs whose spheres touch pass (examples/main.rs:461-479)
    Spheres(&'a [[f32; 4]]),
}

pub struct Layer<Index: SpatialIndex, ID: ObjectID> {
    handle: *mut BpLayer,
    _marker: PhantomData<(Index, ID)>,
}

impl<Index: SpatialIndex, ID: ObjectID> Layer<Index, ID> {
    fn check(&self, status: c_int) {
        if status != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(bp_layer_last_error(self.handle)) };
            panic!("broadphase-b200: status {}: {}", status, msg.to_string_lossy());
        }
    }

    /// src/layer.rs:84-88
    pub fn clear(&mut self) {
        let s = unsafe { bp_layer_clear(self.handle) };
        self.check(s)
    }

    /// src/layer.rs:94-121.  The iterator is collected into two flat arrays (the ABI takes arrays).
    pub fn extend<Iter>(&mut self, system_bounds: Bounds<Index::Point>, objects: Iter)
    where
        Iter: Iterator<Item = (Bounds<Index::Point>, ID)>,
    {
        let (bounds, ids): (Vec<Bounds<Index::Point>>, Vec<ID>) = objects.unzip();
        let s = unsafe {
            bp_layer_extend_host(
                self.handle,
                &system_bounds as *const _ as *const f32,
                bounds.as_ptr() as *const f32,
                ids.as_ptr() as *const c_void,
                ids.len(),
            )
        };
        self.check(s)
    }

    /// src/layer.rs:127-138
    pub fn merge(&mut self, other: &Layer<Index, ID>) {
        let s = unsafe { bp_layer_merge(self.handle, other.handle) };
        self.check(s)
    }

    /// src/layer.rs:157-165
    pub fn sort(&mut self) {
        let s = unsafe { bp_layer_sort(self.handle) };
        self.check(s)
    }

    /// src/layer.rs:146-152
    pub fn par_sort(&mut self) {
        self.sort()
    }

    /// src/layer.rs:456-477.  The slice borrows the layer's pinned result buffer until the next call.
    pub fn scan_filtered<'a>(&'a mut self, filter: Filter) -> &'a [(ID, ID)] {
        let f = match filter {
            Filter::None => BpFilter { kind: 0, table_on_device: 0, arg: 0, table: std::ptr::null(), n_table: 0 },
            Filter::IdParity => BpFilter { kind: 1, table_on_device: 0, arg: 0, table: std::ptr::null(), n_table: 0 },
            Filter::XorMask(m) => BpFilter { kind: 2, table_on_device: 0, arg: m, table: std::ptr::null(), n_table: 0 },
            Filter::Category(t) => BpFilter { kind: 3, table_on_device: 0, arg: 0, table: t.as_ptr() as *const u32, n_table: t.len() },
            Filter::Spheres(t) => BpFilter { kind: 4, table_on_device: 0, arg: 0, table: t.as_ptr() as *const u32, n_table: t.len() },
        };
        let mut pairs: *const c_void = std::ptr::null();
        let mut n: usize = 0;
        let s = unsafe { bp_layer_scan(self.handle, &f, &mut pairs, &mut n) };
        self.check(s);
        if n == 0 { &[] } else { unsafe { std::slice::from_raw_parts(pairs as *const (ID, ID), n) } }
    }

    /// src/layer.rs:449-453
    pub fn scan<'a>(&'a mut self) -> &'a [(ID, ID)] {
        self.scan_filtered(Filter::None)
    }

    /// src/layer.rs:482-487
    pub fn par_scan<'a>(&'a mut self) -> &'a [(ID, ID)] {
        self.scan_filtered(Filter::None)
    }

    /// src/layer.rs:489-520
    pub fn par_scan_filtered<'a>(&'a mut self, filter: Filter) -> &'a [(ID, ID)] {
        self.scan_filtered(filter)
    }

    /// src/layer.rs:293-311, for a batch of boxes: `ids[offsets[q]..offsets[q + 1]]` is what the reference's
    /// `test_box(system_bounds, boxes[q], max_depth)` returns (sorted, duplicate-free).
    pub fn test_box_batch(&mut self, system_bounds: Bounds<Index::Point>, boxes: &[Bounds<Index::Point>], max_depth: Option<u32>)
        -> (Vec<u32>, Vec<ID>)
    {
        let (mut pairs, mut offsets): (*const c_void, *const u32) = (std::ptr::null(), std::ptr::null());
        let mut n: usize = 0;
        let s = unsafe {
            bp_layer_test_box_batch(self.handle, &system_bounds as *const _ as *const f32, boxes.as_ptr() as *const f32, boxes.len(),
                                    max_depth.map_or(-1, |d| d as i32), 0, &mut pairs, &mut offsets, &mut n)
        };
        self.check(s);
        let off = unsafe { std::slice::from_raw_parts(offsets, boxes.len() + 1) }.to_vec();
        let ids = if n == 0 { Vec::new() } else {
            unsafe { std::slice::from_raw_parts(pairs as *const (ID, ID), n) }.iter().map(|&(_query, id)| id).collect()
        };
        (off, ids)
    }

    /// src/layer.rs:293-311
    pub fn test_box(&mut self, system_bounds: Bounds<Index::Point>, test_bounds: Bounds<Index::Point>, max_depth: Option<u32>) -> Vec<ID> {
        self.test_box_batch(system_bounds, &[test_bounds], max_depth).1
    }

    /// src/layer.rs:326-351, for a batch of rays given as rows of `2 * DIM + 2` floats
    /// (origin.., direction.., range_min, range_max).
    pub fn test_ray_batch(&mut self, system_bounds: Bounds<Index::Point>, rays: &[f32], max_depth: Option<u32>) -> (Vec<u32>, Vec<ID>) {
        let nq = rays.len() / (2 * Index::DIM + 2);
        let (mut pairs, mut offsets): (*const c_void, *const u32) = (std::ptr::null(), std::ptr::null());
        let mut n: usize = 0;
        let s = unsafe {
            bp_layer_test_ray_batch(self.handle, &system_bounds as *const _ as *const f32, rays.as_ptr(), nq,
                                    max_depth.map_or(-1, |d| d as i32), 0, &mut pairs, &mut offsets, &mut n)
        };
        self.check(s);
        let off = unsafe { std::slice::from_raw_parts(offsets, nq + 1) }.to_vec();
        let ids = if n == 0 { Vec::new() } else {
            unsafe { std::slice::from_raw_parts(pairs as *const (ID, ID), n) }.iter().map(|&(_query, id)| id).collect()
        };
        (off, ids)
    }

    /// src/layer.rs:424-446 for a batch of rays (rows of `2 * DIM` floats: origin.., direction..): per ray
    /// `Some((dist, id, point))` of the nearest object within `max_dist`, or `None`.
    pub fn pick_ray_batch(&mut self, system_bounds: Bounds<Index::Point>, rays: &[f32], max_dist: f32, max_depth: Option<u32>,
                          shapes: PickShapes) -> Vec<Option<(f32, u64, [f32; 3])>> {
        let nq = rays.len() / (2 * Index::DIM);
        let (kind, table, width) = match shapes {
            PickShapes::Spheres(t) => (0, t, Index::DIM + 1),
            PickShapes::Aabbs(t) => (1, t, 2 * Index::DIM),
        };
        let mut res: *const BpPickResult = std::ptr::null();
        let s = unsafe {
            bp_layer_pick_ray_batch(self.handle, &system_bounds as *const _ as *const f32, rays.as_ptr(), nq, max_dist,
                                    max_depth.map_or(-1, |d| d as i32), kind, table.as_ptr(), table.len() / width, 0, &mut res)
        };
        self.check(s);
        unsafe { std::slice::from_raw_parts(res, nq) }.iter().map(|r| if r.hit != 0 { Some((r.dist, r.id, r.point)) } else { None }).collect()
    }

    /// src/layer.rs:79-81 (copies the tree back from the device)
    pub fn iter(&mut self) -> Vec<(Index::Key, ID)> {
        let (mut k, mut i): (*const c_void, *const c_void) = (std::ptr::null(), std::ptr::null());
        let (mut n, mut sorted): (usize, c_int) = (0, 0);
        let s = unsafe { bp_layer_records(self.handle, &mut k, &mut i, &mut n, &mut sorted) };
        self.check(s);
        let keys = unsafe { std::slice::from_raw_parts(k as *const Index::Key, n) };
        let ids = unsafe { std::slice::from_raw_parts(i as *const ID, n) };
        keys.iter().cloned().zip(ids.iter().cloned()).collect()
    }
}

impl<Index: SpatialIndex, ID: ObjectID> Layer<Index, ID> {
    /// The host mirror of the tree (valid until the next call on this layer) and its sorted flag: what `Clone` and
    /// `PartialEq` read.
    fn raw_records(&self) -> (*const c_void, *const c_void, usize, bool) {
        let (mut k, mut i): (*const c_void, *const c_void) = (std::ptr::null(), std::ptr::null());
        let (mut n, mut sorted): (usize, c_int) = (0, 0);
        let s = unsafe { bp_layer_records(self.handle, &mut k, &mut i, &mut n, &mut sorted) };
        self.check(s);
        (k, i, n, sorted != 0)
    }

    fn current_min_depth(&self) -> u32 {
        let mut d: u32 = 0;
        let s = unsafe { bp_layer_min_depth(self.handle, &mut d) };
        self.check(s);
        d
    }
}

/// src/layer.rs:576-587: `min_depth` and `tree` -- the `(Index, ID)` sequence AND its sorted flag
impl<Index: SpatialIndex, ID: ObjectID> PartialEq for Layer<Index, ID> {
    fn eq(&self, other: &Self) -> bool {
        if self.current_min_depth() != other.current_min_depth() {
            return false;
        }
        let (ka, ia, na, sa) = self.raw_records();
        let (kb, ib, nb, sb) = other.raw_records();
        if na != nb || sa != sb {
            return false;
        }
        let (ka, kb) = unsafe { (std::slice::from_raw_parts(ka as *const Index::Key, na), std::slice::from_raw_parts(kb as *const Index::Key, nb)) };
        let (ia, ib) = unsafe { (std::slice::from_raw_parts(ia as *const ID, na), std::slice::from_raw_parts(ib as *const ID, nb)) };
        ka == kb && ia == ib
    }
}
impl<Index: SpatialIndex, ID: ObjectID> Eq for Layer<Index, ID> {}

/// src/layer.rs:597-617: `min_depth` and the tree with its sorted flag; the copy's result buffers start empty
impl<Index: SpatialIndex, ID: ObjectID> Clone for Layer<Index, ID> {
    fn clone(&self) -> Self {
        let cfg = BpLayerConfig {
            index_kind: Index::KIND,
            id_bytes: ID::BYTES,
            min_depth: self.current_min_depth(),
            device: -1,
            index_capacity: 0,
            collision_capacity: 0,
            test_capacity: 0,
        };
        let mut handle: *mut BpLayer = std::ptr::null_mut();
        let s = unsafe { bp_layer_create(&cfg, &mut handle) };
        assert!(s == 0, "bp_layer_create failed with status {}", s);
        let out = Layer { handle, _marker: PhantomData };
        let (k, i, n, sorted) = self.raw_records();
        let s = unsafe { bp_layer_set_records(out.handle, k, i, n, sorted as c_int, 0) };
        out.check(s);
        out
    }
}

impl<Index: SpatialIndex, ID: ObjectID> Drop for Layer<Index, ID> {
    fn drop(&mut self) {
        unsafe { bp_layer_destroy(self.handle); }
    }
}

/// src/layer.rs:620-696
#[derive(Default)]
pub struct LayerBuilder {
    min_depth: u32,
    index_capacity: Option<usize>,
    collision_capacity: Option<usize>,
    test_capacity: Option<usize>,
}

impl LayerBuilder {
    pub fn new() -> Self { Self::default() }
    pub fn with_min_depth(&mut self, depth: u32) -> &mut Self { self.min_depth = depth; self }
    pub fn with_index_capacity(&mut self, capacity: usize) -> &mut Self { self.index_capacity = Some(capacity); self }
    pub fn with_collision_capacity(&mut self, capacity: usize) -> &mut Self { self.collision_capacity = Some(capacity); self }
    pub fn with_test_capacity(&mut self, capacity: usize) -> &mut Self { self.test_capacity = Some(capacity); self }
    pub fn build<Index: SpatialIndex, ID: ObjectID>(&self) -> Layer<Index, ID> {
        let cfg = BpLayerConfig {
            index_kind: Index::KIND,
            id_bytes: ID::BYTES,
            min_depth: self.min_depth,
            device: -1,
            index_capacity: self.index_capacity.unwrap_or(0),
            collision_capacity: self.collision_capacity.unwrap_or(0),
            test_capacity: self.test_capacity.unwrap_or(0),
        };
        let mut handle: *mut BpLayer = std::ptr::null_mut();
        let s = unsafe { bp_layer_create(&cfg, &mut handle) };
        assert!(s == 0, "bp_layer_create failed with status {} (no CUDA device? there is no CPU fallback)", s);
        Layer { handle, _marker: PhantomData }
    }
}


// ---- several GPUs: the sharded frame (include/bp.h "bp_dist_*"; no counterpart in the reference crate) -------------------------
#[repr(C)]
pub struct BpDist {
    _private: [u8; 0],
}
#[repr(C)]
pub struct BpDistConfig {
    pub index_kind: i32,
    pub min_depth: u32,
    pub device: i32,
    pub rank: i32,
    pub world: i32,
    pub record_capacity: usize,
    pub pair_capacity: usize,
}
extern "C" {
    pub fn bp_dist_create(cfg: *const BpDistConfig, out: *mut *mut BpDist) -> c_int;
    pub fn bp_dist_destroy(ctx: *mut BpDist) -> c_int;
    pub fn bp_dist_handle_bytes() -> usize;
    pub fn bp_dist_export(ctx: *mut BpDist, out_blob: *mut u8) -> c_int;
    pub fn bp_dist_connect(ctx: *mut BpDist, all_blobs: *const u8) -> c_int;
    pub fn bp_dist_set_static(ctx: *mut BpDist, system_bounds: *const f32, d_bounds: *const f32, d_ids: *const c_void, n: usize) -> c_int;
    pub fn bp_dist_frame(ctx: *mut BpDist, system_bounds: *const f32, d_bounds: *const f32, d_ids: *const c_void, n: usize,
                         filter: *const BpFilter, out_d_pairs: *mut *const c_void, out_count: *mut usize) -> c_int;
}
