"""The Rust shim (broadphase-rs_b200/rust/src/lib.rs) is source only -- no cargo / rustc in the build image -- so nothing
compiles it against include/bp.h.  This keeps the two from drifting apart: every `extern "C"` item of the shim must be a
function the header declares, with the same number of parameters and the same machine type in every position, and every
`#[repr(C)]` struct that crosses the boundary must list the header's fields, in order, with matching types."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_SCALARS = {"int": "i32", "int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "size_t": "usize", "float": "f32",
             "double": "f64", "char": "c_char"}
RUST_SCALARS = {"c_int": "i32", "i32": "i32", "u32": "u32", "u64": "u64", "i64": "i64", "usize": "usize", "f32": "f32", "f64": "f64",
                "c_char": "c_char", "u8": "u8"}
STRUCTS = {"bp_layer": "BpLayer", "bp_layer_config": "BpLayerConfig", "bp_filter": "BpFilter", "bp_pick_result": "BpPickResult",
           "bp_dist": "BpDist", "bp_dist_config": "BpDistConfig"}


def _strip_c(text):
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def _strip_rust(text):
    return re.sub(r"//[^\n]*", "", text)


def _c_type(decl):
    """'const float *system_bounds' -> ('f32', 1): pointee machine type and pointer depth (void / opaque blobs -> 'void')."""
    decl = decl.strip()
    depth = decl.count("*")
    words = [w for w in re.sub(r"[*]", " ", decl).split() if w not in ("const", "struct", "unsigned")]
    base = words[0]
    if base == "void":
        return "void", depth
    if base in STRUCTS:
        return STRUCTS[base], depth
    return C_SCALARS.get(base, "struct " + base), depth  # (structs the shim does not use: never compared)


def _rust_type(ty):
    ty = ty.strip()
    depth = 0
    while True:
        m = re.match(r"\*(const|mut)\s+(.*)", ty)
        if not m:
            break
        depth += 1
        ty = m.group(2).strip()
    if ty == "c_void":
        return "void", depth
    if ty in STRUCTS.values():
        return ty, depth
    return RUST_SCALARS[ty], depth


def _same(c, r):
    # an opaque byte blob is `void *` in the header and `*mut u8` / `*const u8` in the shim
    if c == r:
        return True
    return c[1] == r[1] and c[1] >= 1 and {c[0], r[0]} == {"void", "u8"}


def _c_functions():
    text = _strip_c(open(os.path.join(ROOT, "include", "bp.h")).read())
    out = {}
    for m in re.finditer(r"\b([a-z_0-9]+(?:\s+[a-z_0-9]+)*\s*\**)\s*\b(bp_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = [] if args in ("", "void") else [a for a in args.split(",")]
        out[name] = (_c_type(ret + " x") if "*" in ret or ret.strip() not in ("int",) else ("i32", 0), [_c_type(p) for p in params])
    return out


def _rust_functions():
    text = _strip_rust(open(os.path.join(ROOT, "broadphase-rs_b200", "rust", "src", "lib.rs")).read())
    out = {}
    for block in re.finditer(r'extern\s+"C"\s*\{(.*?)\n\}', text, flags=re.S):
        for m in re.finditer(r"fn\s+(bp_[a-z_0-9]+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block.group(1), flags=re.S):
            name, args, ret = m.group(1), m.group(2).strip(), (m.group(3) or "").strip()
            params = []
            for a in [a for a in args.split(",") if a.strip()]:
                params.append(_rust_type(a.split(":", 1)[1]))
            out[name] = (_rust_type(ret) if ret else None, params)
    return out


def _c_struct_fields(name):
    text = _strip_c(open(os.path.join(ROOT, "include", "bp.h")).read())
    body = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), text, flags=re.S).group(1)
    fields = []
    for stmt in [s.strip() for s in body.split(";") if s.strip()]:
        m = re.match(r"(.*?)([a-z_0-9]+(?:\s*,\s*[a-z_0-9]+)*)\s*(\[\d+\])?$", stmt)
        ty, names, arr = m.group(1), m.group(2), m.group(3)
        for n in [n.strip() for n in names.split(",")]:
            fields.append((n, _c_type(ty + " x") + ((int(arr[1:-1]),) if arr else ())))
    return fields


def _rust_struct_fields(name):
    text = _strip_rust(open(os.path.join(ROOT, "broadphase-rs_b200", "rust", "src", "lib.rs")).read())
    body = re.search(r"pub\s+struct\s+%s\s*\{(.*?)\}" % name, text, flags=re.S).group(1)
    fields = []
    for m in re.finditer(r"pub\s+([a-z_0-9]+)\s*:\s*([^,\n]+)", body):
        ty = m.group(2).strip()
        arr = re.match(r"\[(.*);\s*(\d+)\]", ty)
        fields.append((m.group(1), _rust_type(arr.group(1)) + (int(arr.group(2)),) if arr else _rust_type(ty)))
    return fields


def test_every_extern_item_of_the_shim_matches_the_header():
    c, r = _c_functions(), _rust_functions()
    assert len(r) >= 19, sorted(r)
    # the Layer API of the crate and the sharded frame are all bound
    for must in ("bp_layer_create", "bp_layer_destroy", "bp_layer_clear", "bp_layer_extend_host", "bp_layer_merge", "bp_layer_sort",
                 "bp_layer_scan", "bp_layer_records", "bp_layer_test_box_batch", "bp_layer_test_ray_batch", "bp_layer_pick_ray_batch",
                 "bp_dist_create", "bp_dist_connect", "bp_dist_export", "bp_dist_frame", "bp_dist_set_static", "bp_dist_destroy"):
        assert must in r, must
    for name, (ret, params) in r.items():
        assert name in c, "%s is not declared in include/bp.h" % name
        c_ret, c_params = c[name]
        assert len(params) == len(c_params), (name, len(params), len(c_params))
        for i, (cp, rp) in enumerate(zip(c_params, params)):
            assert _same(cp, rp), "%s: parameter %d is %s in bp.h and %s in lib.rs" % (name, i, cp, rp)
        assert ret is not None and _same(c_ret, ret), (name, c_ret, ret)


def test_structs_that_cross_the_boundary_have_the_headers_layout():
    for c_name in ("bp_layer_config", "bp_filter", "bp_pick_result", "bp_dist_config"):
        cf, rf = _c_struct_fields(c_name), _rust_struct_fields(STRUCTS[c_name])
        assert [n for n, _ in cf] == [n for n, _ in rf], (c_name, cf, rf)
        for (n, ct), (_, rt) in zip(cf, rf):
            assert ct == rt, "%s.%s is %s in bp.h and %s in lib.rs" % (c_name, n, ct, rt)
