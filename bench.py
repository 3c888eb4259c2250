#!/usr/bin/env python
"""bench.py -- objects/s (and pairs/s) for extend + sort + scan, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg5|cfg3|cfg4|cfg2|cfg1]

A "step" is one frame of the hot path over one batch of synthetic AABBs:
    clear -> extend -> par_sort -> [merge(static)] -> par_scan(_filtered)
Headline workload at every N (task 2): the per-GPU shape of BASELINE config 5 -- 2^25 uniform-size
AABBs per GPU (N = 8: 2^28 objects), Index64_3D.  The 16M-object target config 3 (log-normal sizes, ID-parity filter)
and config 4 (2^26 static + 2^22 dynamic objects per frame through Layer::merge) are measured in the same run and
reported under `other_workloads`, EACH with its own `e2e`, `cpu_baseline` and `roofline`; configs 2 and 1 ride along
as short blocks.

`value`   : whole-job objects/s with the AABBs already resident in HBM and the pairs left in HBM, timed per step with
            CUDA events on the layer's stream (256 MiB written between steps: L2 flushed; the inputs exceed L2 anyway).
`e2e`     : the same frame through the host-buffer C-ABI calls (bp_layer_extend_host + bp_layer_scan): H2D of the
            AABBs / IDs from pinned memory and D2H of the pair list inside the timed region.
`roofline`: the dominant kernel class, algorithmic bytes / CUDA-event time (per-launch event pairs, profiling mode of
            the library), against the measured HBM peak; `traffic` = DRAM bytes per launch from the committed ncu capture.
`cpu_baseline` / --impl reference: the reference's frame on the host cores.  The Rust crate cannot be built here (no
            cargo / rustc), so this is the C++ restatement in oracle/ (kind "port" -- pinned by the reference's golden
            files, tests/test_reference_fixtures.py): sequential extend + parallel sort + par_scan, all host threads.
            The reference arm never imports the product package (scenes are loaded by path, numpy only).
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "objects/sec for extend+sort+scan"
DTYPE = "u64 keys / u32 ids (f32 quantiser)"
L2_NOTE = ("GPU arm: 256 MiB buffer written between steps, outside the per-step CUDA-event windows (the inputs also exceed "
           "the 126 MB L2); CPU arm: not applicable")

# name -> description, objects per step (per GPU for cfg5), filter, CPU-sample shift (the CPU baseline runs the same
# recipe at n / 8^shift objects: the cube edge follows n^(-1/3), so a factor 8 in n is exactly one octree level and the
# sample has the same records / object and pairs / object as the full scene)
WORKLOADS = {
    "cfg5": dict(desc="2^25 uniform-size AABBs per GPU (edge 0.4*N_total^-1/3; BASELINE config 5's per-GPU shape: 2^28 on 8 GPUs), "
                      "Index64_3D, u32 IDs, extend+par_sort+par_scan", n=1 << 25, parity=False, cpu_shift=1),
    "cfg3": dict(desc="16M (2^24) log-normal AABBs (multi-depth keys), Index64_3D, u32 IDs, scan_filtered ID-parity", n=1 << 24,
                 parity=True, cpu_shift=1),
    "cfg4": dict(desc="64M (2^26) static layer pre-sorted + 4M (2^22) dynamic objects per frame via Layer::merge, Index64_3D, "
                      "u32 IDs; frame = clear+extend+sort(dynamic)+merge(static)+par_scan", n=(1 << 26) + (1 << 22), parity=False,
                 cpu_shift=1, n_static=1 << 26, n_dynamic=1 << 22),
    "cfg2": dict(desc="1M (2^20) uniform-size AABBs (edge 0.4*N^-1/3), Index64_3D, extend+par_sort+par_scan", n=1 << 20,
                 parity=False, cpu_shift=0),
    "cfg1": dict(desc="examples/main.rs scene: 10,000 circles, Index32_2D, min_depth 4, par_scan", n=10_000, parity=False,
                 cpu_shift=0),
}


def load_scenes():
    """broadphase-rs_b200/scenes.py by path: numpy only, loads neither the package nor libbroadphase_b200.so."""
    spec = importlib.util.spec_from_file_location("bp_scenes", os.path.join(ROOT, "broadphase-rs_b200", "scenes.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_inputs(scenes, wl, shift=0, world=1, rank=0, variants=2):
    """The synthetic inputs of one workload at 1 / 8^shift of its size.
    -> dict(frames=[scene, ...] (alternated step by step), static=scene|None, n=objects per step on this rank)."""
    w = WORKLOADS[wl]
    div = 8 ** shift
    if wl == "cfg1":
        return dict(frames=[scenes.example_circles(w["n"], 1 + v) for v in range(variants)], static=None, n=w["n"])
    if wl in ("cfg2", "cfg5"):
        n = w["n"] // div
        n_total = n * world
        ef = 0.4 * (float(n_total) / float(n)) ** (-1.0 / 3.0)     # the edge follows the GLOBAL object count
        seed = 2 if wl == "cfg2" else 6
        return dict(frames=[scenes.uniform_cubes(n, seed + 100 * v + 1000 * rank, id_base=rank * n, edge_factor=ef)
                            for v in range(variants)], static=None, n=n)
    if wl == "cfg3":
        n = w["n"] // div
        return dict(frames=[scenes.lognormal_cubes(n, 3 + 100 * v) for v in range(variants)], static=None, n=n)
    if wl == "cfg4":
        ns, nd = w["n_static"] // div, w["n_dynamic"] // div
        static = scenes.uniform_cubes(ns, 4)
        ef = 0.4 * (ns / nd) ** (-1.0 / 3.0)                     # dynamic objects have the static objects' size
        return dict(frames=[scenes.uniform_cubes(nd, 5 + v, id_base=ns, edge_factor=ef) for v in range(variants)],
                    static=static, n=ns + nd)
    raise ValueError(wl)


def make_scene(bp_or_scenes, wl, n=None):
    """One scene of a workload's recipe at n objects (tools/: profile_frame.py, timeline_frame.py)."""
    sm = getattr(bp_or_scenes, "scenes", bp_or_scenes)
    n = n or (WORKLOADS[wl]["n_dynamic"] if wl == "cfg4" else WORKLOADS[wl]["n"])
    if wl == "cfg1":
        return sm.example_circles(n, 1)
    if wl == "cfg3":
        return sm.lognormal_cubes(n, 3)
    return sm.uniform_cubes(n, 6 if wl == "cfg5" else 2)


def scene_filter(bp, wl):
    return bp.ScanFilter.id_parity() if WORKLOADS[wl]["parity"] else None


def gpu_frame_device(layer, sc, d_bounds, d_ids, n, flt):
    layer.clear()
    layer.extend_device(sc["sys_bounds"], d_bounds, d_ids, n)
    layer.par_sort()
    return layer.scan_device(flt)


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons),
                "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(wl, kclass):
    """dram bytes per launch of a kernel class at a workload, from the committed ncu capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(wl, {}).get(kclass)
    except Exception:
        return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------
# the reference's frame on the host cores (oracle port)
class CpuFrame:
    def __init__(self, co, wl, inp):
        f0 = inp["frames"][0]
        self.co, self.inp, self.parity = co, inp, WORKLOADS[wl]["parity"]
        self.layer = co.OracleLayer(f0["kind"], 4, f0["min_depth"])
        self.static = None
        if inp["static"] is not None:      # built and sorted once, outside the timed frames (like the GPU arm)
            st = inp["static"]
            self.static = co.OracleLayer(st["kind"], 4, st["min_depth"])
            self.static.extend(st["sys_bounds"], st["bounds"], st["ids"])
            self.static.par_sort()

    def run(self, step):
        sc = self.inp["frames"][step % len(self.inp["frames"])]
        L, co = self.layer, self.co
        t0 = time.perf_counter()
        L.clear()
        L.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
        L.par_sort()
        if self.static is not None:
            L.merge(self.static)           # the reference appends and re-sorts everything (src/layer.rs:127-138)
        co.lib().bpo_layer_par_scan(L._h, co.FILTER_ID_PARITY if self.parity else co.FILTER_NONE, 0, None, 0)
        dt = time.perf_counter() - t0
        return dt, int(co.lib().bpo_layer_num_collisions(L._h))


def cpu_baseline(co, scenes, wl, budget_s=10.0, world=1):
    """cpu_baseline block: the oracle port on a bounded sample of the workload, all host threads."""
    co.lib().bpo_set_threads(host_threads())   # torchrun exports OMP_NUM_THREADS=1 to its workers
    shift = WORKLOADS[wl]["cpu_shift"]
    inp = make_inputs(scenes, wl, shift, world=world)
    fr = CpuFrame(co, wl, inp)
    dt, _ = fr.run(0)                        # warm-up, sizes the loop
    steps = int(max(2, min(100, budget_s / max(dt, 1e-3))))
    times, pairs = [], 0
    for s in range(steps):
        dt, pairs = fr.run(s + 1)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return {"value": inp["n"] / (ms * 1e-3), "unit": "objects/s", "cores": int(co.lib().bpo_max_threads()), "kind": "port",
            "sample": "%d-object frames of the same recipe (1/%d of the workload: one octree level shallower per factor 8, same "
                      "records and pairs per object), %d timed after 1 warm-up; C++ restatement of the reference (no Rust "
                      "toolchain), sequential extend + parallel sort + par_scan" % (inp["n"], 8 ** shift, steps),
            "ms_per_step": ms, "pairs_per_s": pairs / (ms * 1e-3), "objects_per_step": inp["n"]}


def run_reference(args):
    """--impl reference: the reference's own CPU path on the host cores, the headline workload at full size
    (N = 1) or rank 0's per-GPU share of it (N > 1: a bounded sample -- the whole N x 2^25 scene does not fit a run of
    a few minutes on the host).  Never touches the GPU or the product library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_oracle as co
    scenes = load_scenes()
    wl = args.workload
    world = max(1, args.gpus)
    co.lib().bpo_set_threads(host_threads())
    inp = make_inputs(scenes, wl, 0, world=world if wl == "cfg5" else 1)
    fr = CpuFrame(co, wl, inp)
    times, pairs = [], 0
    for s in range(args.warmup + args.steps):
        dt, pairs = fr.run(s)
        if s >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    n = inp["n"]
    value = n / (ms * 1e-3)
    if world > 1:
        sample = ("rank 0's share of the %d-GPU scene: %d-object frames (2^25 per GPU at the %d x 2^25 scene's cube edge), %d timed; "
                  "objects/s of this sample" % (world, n, world, len(times)))
    else:
        sample = "the full workload: %d-object frames (clear+extend+par_sort+par_scan), %d timed" % (n, len(times))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "objects/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(wl, world),
        "pairs_per_s": pairs / (ms * 1e-3), "objects_in_sample": n,
        "cpu_baseline": {"value": value, "unit": "objects/s", "cores": int(co.lib().bpo_max_threads()), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "objects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)
    return 0


def workload_config(wl, world=1):
    """`config` of the JSON line -- identical in both arms."""
    w = WORKLOADS[wl]
    n = w["n"] * (world if wl == "cfg5" else 1)
    cfg = {"workload": w["desc"], "objects_per_step": n, "index": "Index32_2D" if wl == "cfg1" else "Index64_3D", "ids": "u32",
           "scenes": "2 scenes of the recipe (different seeds) alternate step by step", "l2": L2_NOTE}
    if world > 1:
        cfg["parallelism"] = "morton-range-shard x%d" % world
        cfg["objects_per_gpu"] = w["n"]
    return cfg


# ---------------------------------------------------------------------------------------------------
class GpuFrame:
    """One workload on one GPU through the C ABI: device-resident and host-buffer variants of the same frame."""

    def __init__(self, bp, torch, wl, inp, device=0):
        self.bp, self.torch, self.wl, self.inp, self.device = bp, torch, wl, inp, device
        f0 = inp["frames"][0]
        self.flt = bp.ScanFilter.id_parity() if WORKLOADS[wl]["parity"] else None
        self.stream = torch.cuda.current_stream(device)
        mk = lambda sc: bp.LayerBuilder().with_min_depth(sc["min_depth"]).with_device(device).build(sc["kind"], "u32")
        self.layer = mk(f0)
        self.layer.set_stream(self.stream.cuda_stream)
        self.static = None
        self.static_ms = None
        if inp["static"] is not None:
            st = inp["static"]
            self.static = mk(st)
            self.static.set_stream(self.stream.cuda_stream)
            sb = torch.from_numpy(st["bounds"]).cuda(device)
            si = torch.from_numpy(st["ids"].view(np.int32)).cuda(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            self.static.extend_device(st["sys_bounds"], sb, si, st["bounds"].shape[0])
            self.static.sort()
            e1.record(self.stream)
            torch.cuda.synchronize(device)
            self.static_ms = e0.elapsed_time(e1)
            del sb, si
        self.d_in = [(torch.from_numpy(sc["bounds"]).cuda(device), torch.from_numpy(sc["ids"].view(np.int32)).cuda(device))
                     for sc in inp["frames"]]
        self.h_in = None
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % device)  # > 126 MB L2

    def layers(self):
        return [l for l in (self.layer, self.static) if l is not None]

    def launches(self):
        return sum(l.stats()["launches_total"] for l in self.layers())

    def frame_device(self, step):
        k = step % len(self.d_in)
        sc, (db, di) = self.inp["frames"][k], self.d_in[k]
        L = self.layer
        L.clear()
        L.extend_device(sc["sys_bounds"], db, di, sc["bounds"].shape[0])
        L.par_sort()
        if self.static is not None:
            L.merge(self.static)
        return L.scan_device(self.flt)

    def frame_host(self, step):
        if self.h_in is None:
            t = self.torch
            self.h_in = [(t.from_numpy(sc["bounds"]).pin_memory().numpy(),
                          t.from_numpy(sc["ids"].view(np.int32)).pin_memory().numpy().view(np.uint32)) for sc in self.inp["frames"]]
        k = step % len(self.h_in)
        sc, (hb, hi) = self.inp["frames"][k], self.h_in[k]
        L = self.layer
        L.clear()
        L.extend(sc["sys_bounds"], hb, hi)
        L.par_sort()
        if self.static is not None:
            L.merge(self.static)
        p = L.par_scan_filtered(self.flt)
        return p, hb.nbytes + hi.nbytes

    def close(self):
        for l in self.layers():
            l.close()
        self.d_in = self.h_in = self.flush = None


def measure_gpu(bp, torch, scenes, wl, steps, warmup, device=0, with_e2e=True, with_profile=True):
    """Times one workload on one GPU.  -> dict of measurements (ms, e2e, per-class profile, counts)."""
    inp = make_inputs(scenes, wl)
    g = GpuFrame(bp, torch, wl, inp, device)
    n, stream, flush = inp["n"], g.stream, g.flush
    out = {"n": n, "static_build_ms": g.static_ms}
    for s in range(warmup):
        g.frame_device(s)
    torch.cuda.synchronize(device)
    for l in g.layers():
        l.reset_stats()
    launches0 = g.launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    pairs = 0
    for s in range(steps):
        flush.fill_(s & 0xFF)
        ev[s][0].record(stream)
        _, pairs = g.frame_device(warmup + s)
        ev[s][1].record(stream)
    torch.cuda.synchronize(device)
    times = [a.elapsed_time(b) for a, b in ev]
    st = g.layer.stats()
    out.update(pairs=pairs, ms_steps=times, ms=sum(times) / len(times), stats=st, launches=g.launches() - launches0)

    if with_profile:  # per-kernel-class CUDA-event timing, same frames, separate loop
        for l in g.layers():
            l.set_profiling(True)
            l.reset_stats()
        for s in range(steps):
            flush.fill_(s & 0xFF)
            g.frame_device(warmup + s)
        torch.cuda.synchronize(device)
        prof = {"kernel_ms": {}, "launches": {}, "algo_bytes": {}}
        for l in g.layers():
            stl = l.stats()
            for k in prof:
                for c, v in stl[k].items():
                    prof[k][c] = prof[k].get(c, 0) + v
            l.set_profiling(False)
        out["profile"] = prof

    if with_e2e:  # host buffers in, host pair list out, through the C ABI
        e2e, h2d, d2h, check = [], 0, 0, 0
        for s in range(warmup + steps):
            flush.fill_(s & 0xFF)
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            p, h2d = g.frame_host(s)
            check = int(p[-1, 0]) if p.shape[0] else 0  # touch the host result
            dt = time.perf_counter() - t0
            d2h = int(p.nbytes)
            if s >= warmup:
                e2e.append(dt * 1e3)
        out.update(e2e_ms=sum(e2e) / len(e2e), h2d=h2d, d2h=d2h, e2e_check=check)
    g.close()
    del g
    torch.cuda.empty_cache()
    return out


def roofline_block(wl, prof, steps, peak, peak_src):
    kclass = max((c for c in prof["kernel_ms"] if c != "misc"), key=lambda c: prof["kernel_ms"][c])  # "misc" is a grab-bag of small helpers, not one kernel
    k_ms, k_launch, k_bytes = prof["kernel_ms"][kclass], prof["launches"][kclass], prof["algo_bytes"][kclass]
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    total = sum(prof["kernel_ms"].values())
    return {"bound": "hbm", "kernel": kclass, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(wl, kclass), "peak_source": peak_src,
            "launches_per_step": k_launch / steps, "avg_launch_ms": k_ms / max(k_launch, 1),
            "algo_bytes_per_launch": k_bytes / max(k_launch, 1),
            "share_of_kernel_time": k_ms / total if total else None,
            "per_class_ms_per_step": {c: v / steps for c, v in prof["kernel_ms"].items() if v > 0},
            "per_class_gbs": {c: prof["algo_bytes"][c] / (v * 1e-3) / 1e9 for c, v in prof["kernel_ms"].items() if v > 0},
            "whole_frame_gbs": sum(prof["algo_bytes"].values()) / (total * 1e-3) / 1e9 if total else None}


def workload_block(wl, m, steps, peak, peak_src, cpu):
    """The full measurement block of one workload (also what `other_workloads` holds)."""
    n, ms, st = m["n"], m["ms"], m["stats"]
    b = {"workload": WORKLOADS[wl]["desc"], "objects_per_step": n, "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "objects/s",
         "pairs_per_s": m["pairs"] / (ms * 1e-3), "steps": steps,
         "counts": {"records": st["n_records"], "work_items": st["n_work_items"], "raw_pairs": st["n_raw_pairs"],
                    "unique_pairs": m["pairs"], "sort_passes": st["sort_passes"], "pair_sort_passes": st["pair_sort_passes"],
                    "merged": st["merged"]},
         "gpu_launches": m["launches"]}
    if m.get("static_build_ms") is not None:
        b["static_build_ms"] = m["static_build_ms"]
        b["dynamic_objects_per_s"] = WORKLOADS[wl]["n_dynamic"] / (ms * 1e-3)
    if "e2e_ms" in m:
        b["e2e"] = {"value": n / (m["e2e_ms"] * 1e-3), "unit": "objects/s", "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["e2e_ms"], "pairs_per_s": m["pairs"] / (m["e2e_ms"] * 1e-3)}
    if "profile" in m:
        b["roofline"] = roofline_block(wl, m["profile"], steps, peak, peak_src)
    if cpu is not None:
        b["cpu_baseline"] = cpu
    return b


def run_ours(args):
    import torch
    import _loadpkg
    bp = _loadpkg.load()
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from broadphase_rs_b200 import dist_bench
        return dist_bench.run(args, bp)
    from oracle import cpu_oracle as co      # cpu_baseline leg only (the checker, never the thing measured)
    scenes = load_scenes()
    wl = args.workload
    torch.cuda.set_device(0)
    peak, peak_src = hbm_peak()
    with ClockSampler(0) as clocks:
        m = measure_gpu(bp, torch, scenes, wl, args.steps, args.warmup)
    head = workload_block(wl, m, args.steps, peak, peak_src, cpu_baseline(co, scenes, wl, args.cpu_budget))

    extra = {}
    if not args.no_extra:
        for other, st_, wu in (("cfg3", 10, 3), ("cfg4", 5, 3), ("cfg2", 20, 3), ("cfg1", 20, 3)):
            if other == wl:
                continue
            try:   # never lose the headline line to an extra measurement
                mo = measure_gpu(bp, torch, scenes, other, st_, wu)
                full = other in ("cfg3", "cfg4")
                extra[other] = workload_block(other, mo, st_, peak, peak_src,
                                              cpu_baseline(co, scenes, other, args.cpu_budget if full else 3.0))
            except Exception as e:
                extra[other + "_error"] = repr(e)

    line = {
        "metric": METRIC, "value": head["value"], "unit": "objects/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(wl),
        "pairs_per_s": head["pairs_per_s"], "counts": head["counts"],
        "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
        "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"], "clocks": clocks.summary(),
    }
    if extra:
        line["other_workloads"] = extra
    emit_line(line)
    return 0


_REAL_STDOUT = None


def emit_line(line):
    """The one JSON line goes to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    # (under torchrun this file runs as __main__ while dist_bench imports it again as `bench`: the second copy of
    # the module finds the descriptor in the environment)
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else int(os.environ.get("BP_BENCH_STDOUT_FD", "-1"))
    if fd >= 0:
        os.write(fd, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    # stdout must carry exactly one JSON line, but libraries print there too (NCCL's version banner,
    # for one): point fd 1 at stderr for the whole run and keep the original for the JSON line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.environ["BP_BENCH_STDOUT_FD"] = str(_REAL_STDOUT)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the other workloads (configs 3, 4, 2, 1)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the untimed oracle / hash self-checks")
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds of CPU work per cpu_baseline block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
