#!/usr/bin/env python
"""bench.py -- objects/s (and pairs/s) for extend + sort + scan, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3]

A "step" is one frame of the hot path over one batch of synthetic AABBs:
    clear -> extend -> par_sort -> par_scan(_filtered)
`value`   : whole-job objects/s with the AABBs already resident in HBM and the pairs left in HBM,
            timed per step with CUDA events on the layer's stream (L2 flushed between steps).
`e2e`     : the same frame through the host-buffer C-ABI calls (bp_layer_extend_host + bp_layer_scan):
            H2D of the AABBs/IDs from pinned memory and D2H of the pair list inside the timed region.
`roofline`: the dominant kernel class, algorithmic bytes / CUDA-event time, against the measured HBM peak.
`cpu_baseline` / --impl reference: the C++ restatement of the reference (oracle/, Rust toolchain is not
            available) on the host cores: sequential extend + parallel sort + par_scan.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg1": dict(desc="examples/main.rs scene: 10,000 circles, Index32_2D, min_depth 4, par_scan", n=10_000),
    "cfg2": dict(desc="1M (2^20) uniform-size AABBs (edge 0.4*N^-1/3), Index64_3D, extend+par_sort+par_scan", n=1 << 20),
    "cfg3": dict(desc="16M (2^24) log-normal AABBs (multi-depth keys), Index64_3D, scan_filtered ID-parity", n=1 << 24),
}


def make_scene(bp, workload, n=None, seed=None, id_base=0):
    w = WORKLOADS[workload]
    n = n or w["n"]
    if workload == "cfg1":
        return bp.scenes.example_circles(n, seed or 1)
    if workload == "cfg2":
        return bp.scenes.uniform_cubes(n, seed or 2, id_base=id_base)
    if workload == "cfg3":
        return bp.scenes.lognormal_cubes(n, seed or 3)
    raise ValueError(workload)


def scene_filter(bp, workload):
    return bp.ScanFilter.id_parity() if workload == "cfg3" else None


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


def ncu_traffic(kclass):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kclass)
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------
def cpu_frames(co, sc, filt, steps, warmup, threads=None):
    """The reference's frame on the host cores (oracle port): clear, extend, par_sort, par_scan."""
    if threads:
        co.lib().bpo_set_threads(threads)
    L = co.OracleLayer(sc["kind"], 4, sc["min_depth"])
    times, pairs = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        L.clear()
        L.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
        L.par_sort()
        if filt:
            co.lib().bpo_layer_par_scan(L._h, co.FILTER_ID_PARITY, 0, None, 0)
        else:
            co.lib().bpo_layer_par_scan(L._h, co.FILTER_NONE, 0, None, 0)
        dt = time.perf_counter() - t0
        pairs = co.lib().bpo_layer_num_collisions(L._h)
        if it >= warmup:
            times.append(dt)
    return times, pairs


def run_reference(args):
    """--impl reference: the reference's own CPU path.  The Rust crate cannot be built in this image
    (no cargo/rustc), so this is the C++ restatement in oracle/ ("port"), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import _loadpkg
    bp = _loadpkg.load()
    from oracle import cpu_oracle as co
    wl = args.workload
    n = WORKLOADS[wl]["n"]
    sample_n = min(n, 1 << 20)  # bounded sample: at most 2^20 objects per step
    sc = make_scene(bp, wl, sample_n)
    # all the host threads there are: torchrun exports OMP_NUM_THREADS=1 to its workers, which would make this a
    # single-threaded run at N > 1
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    co.lib().bpo_set_threads(avail)
    cores = co.lib().bpo_max_threads()
    times, pairs = cpu_frames(co, sc, wl == "cfg3", args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    value = sample_n / (ms * 1e-3)
    sample = "%d-object frames of the same recipe (clear+extend+par_sort+par_scan), %d timed" % (sample_n, len(times))
    line = {
        "impl": "reference", "metric": "objects/sec for extend+sort+scan", "value": value, "unit": "objects/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 keys / u32 ids (f32 quantiser)", "data": "synthetic",
        "config": {"workload": WORKLOADS[wl]["desc"], "objects_per_step": sample_n},
        "pairs_per_s": pairs / (ms * 1e-3),
        "cpu_baseline": {"value": value, "unit": "objects/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "objects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)
    return 0


# ---------------------------------------------------------------------------------------------------
def gpu_frame_device(layer, sc, d_bounds, d_ids, n, flt):
    layer.clear()
    layer.extend_device(sc["sys_bounds"], d_bounds, d_ids, n)
    layer.par_sort()
    return layer.scan_device(flt)


def time_workload(bp, torch, wl, steps, warmup, n=None, with_e2e=True, with_profile=True, device=0):
    """Times one workload on one GPU.  Returns a dict of measurements."""
    sc = make_scene(bp, wl, n)
    n = sc["bounds"].shape[0]
    flt = scene_filter(bp, wl)
    id_np = sc["ids"]
    d_bounds = torch.from_numpy(sc["bounds"]).cuda(device)
    d_ids = torch.from_numpy(id_np.view(np.int32)).cuda(device)
    layer = bp.LayerBuilder().with_min_depth(sc["min_depth"]).with_device(device).build(sc["kind"], "u32")
    stream = torch.cuda.current_stream(device)
    layer.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % device)  # > 126 MB L2

    out = {}
    for _ in range(warmup):
        gpu_frame_device(layer, sc, d_bounds, d_ids, n, flt)
    torch.cuda.synchronize(device)
    layer.reset_stats()
    launches0 = layer.stats()["launches_total"]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    pairs = 0
    for s in range(steps):
        flush.fill_(s & 0xFF)
        ev[s][0].record(stream)
        _, pairs = gpu_frame_device(layer, sc, d_bounds, d_ids, n, flt)
        ev[s][1].record(stream)
    torch.cuda.synchronize(device)
    times = [a.elapsed_time(b) for a, b in ev]
    st = layer.stats()
    out.update(n=n, pairs=pairs, ms_steps=times, ms=sum(times) / len(times), stats=st,
               launches=(st["launches_total"] - launches0))

    if with_profile:  # per-kernel-class CUDA-event timing, same frames, separate loop
        layer.set_profiling(True)
        layer.reset_stats()
        for s in range(steps):
            flush.fill_(s & 0xFF)
            gpu_frame_device(layer, sc, d_bounds, d_ids, n, flt)
        torch.cuda.synchronize(device)
        out["profile"] = layer.stats()
        layer.set_profiling(False)

    if with_e2e:  # host buffers in, host pair list out, through the C ABI
        h_bounds = torch.from_numpy(sc["bounds"]).pin_memory().numpy()
        h_ids = torch.from_numpy(id_np.view(np.int32)).pin_memory().numpy().view(np.uint32)
        e2e = []
        for s in range(warmup + steps):
            flush.fill_(s & 0xFF)
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            layer.clear()
            layer.extend(sc["sys_bounds"], h_bounds, h_ids)
            layer.par_sort()
            p = layer.par_scan_filtered(flt)
            checksum = int(p[-1, 0]) if p.shape[0] else 0  # touch the host result
            dt = time.perf_counter() - t0
            if s >= warmup:
                e2e.append(dt * 1e3)
        out.update(e2e_ms=sum(e2e) / len(e2e), h2d=h_bounds.nbytes + h_ids.nbytes, d2h=int(p.nbytes), e2e_check=checksum)
    layer.close()
    del d_bounds, d_ids, flush
    return out, sc


def run_ours(args):
    import torch
    import _loadpkg
    bp = _loadpkg.load()
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from broadphase_rs_b200 import dist_bench
        return dist_bench.run(args, bp)
    wl = args.workload
    torch.cuda.set_device(0)
    peak, peak_src = hbm_peak()
    with ClockSampler(0) as clocks:
        m, sc = time_workload(bp, torch, wl, args.steps, args.warmup)
    n, ms = m["n"], m["ms"]
    value = n / (ms * 1e-3)

    # roofline of the dominant kernel class
    prof = m["profile"]
    kclass = max((c for c in prof["kernel_ms"] if c != "misc"), key=lambda c: prof["kernel_ms"][c])  # "misc" is a grab-bag of small helpers, not one kernel
    k_ms, k_launch, k_bytes = prof["kernel_ms"][kclass], prof["launches"][kclass], prof["algo_bytes"][kclass]
    achieved = (k_bytes / max(k_launch, 1)) / (k_ms / max(k_launch, 1) * 1e-3) / 1e9 if k_ms > 0 else 0.0
    total_k_ms = sum(prof["kernel_ms"].values())
    roofline = {"bound": "hbm", "kernel": kclass, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(kclass), "peak_source": peak_src,
                "launches_per_step": k_launch / args.steps, "avg_launch_ms": k_ms / max(k_launch, 1),
                "algo_bytes_per_launch": k_bytes / max(k_launch, 1),
                "share_of_kernel_time": k_ms / total_k_ms if total_k_ms else None,
                "per_class_ms_per_step": {c: prof["kernel_ms"][c] / args.steps for c in prof["kernel_ms"]},
                "per_class_gbs": {c: (prof["algo_bytes"][c] / (prof["kernel_ms"][c] * 1e-3) / 1e9) if prof["kernel_ms"][c] > 0 else None
                                  for c in prof["kernel_ms"]}}

    # CPU baseline (oracle port) on a bounded sample, rank 0 only
    from oracle import cpu_oracle as co
    cores = co.lib().bpo_max_threads()
    sample_n = min(n, 1 << 20)
    sc_cpu = sc if sample_n == n else make_scene(bp, wl, sample_n)
    # bounded sample: about 10-15 s of CPU work (one frame first, to size the loop)
    probe, _ = cpu_frames(co, sc_cpu, wl == "cfg3", 1, 1)
    cpu_steps = int(max(3, min(200, 12.0 / max(probe[0], 1e-3))))
    cpu_times, cpu_pairs = cpu_frames(co, sc_cpu, wl == "cfg3", cpu_steps, 0)
    cpu_ms = 1e3 * sum(cpu_times) / len(cpu_times)
    cpu_baseline = {"value": sample_n / (cpu_ms * 1e-3), "unit": "objects/s", "cores": cores, "kind": "port",
                    "sample": "%d-object frames (clear+extend+par_sort+par_scan), %d timed after 2 warm-ups; C++ restatement "
                              "of the reference (no Rust toolchain)" % (sample_n, cpu_steps),
                    "ms_per_step": cpu_ms}

    extra = {}
    if wl == "cfg2" and not args.no_extra:  # the 16M-object target of BASELINE.json's north_star, reported beside it
        try:
            m3, _ = time_workload(bp, torch, "cfg3", max(3, min(args.steps, 5)), 3, with_e2e=False, with_profile=True)
            p3 = m3["profile"]
            extra["cfg3_16M_lognormal_parity_filter"] = {
                "objects": m3["n"], "ms_per_step": m3["ms"], "objects_per_s": m3["n"] / (m3["ms"] * 1e-3),
                "pairs": m3["pairs"], "records": m3["stats"]["n_records"], "raw_pairs": m3["stats"]["n_raw_pairs"],
                "sort_passes": m3["stats"]["sort_passes"], "pair_sort_passes": m3["stats"]["pair_sort_passes"],
                "per_class_ms_per_step": {c: p3["kernel_ms"][c] / max(3, min(args.steps, 5)) for c in p3["kernel_ms"]},
                "sort_pass_gbs": (p3["algo_bytes"]["sort_pass"] / (p3["kernel_ms"]["sort_pass"] * 1e-3) / 1e9)
                if p3["kernel_ms"]["sort_pass"] > 0 else None,
                "sort_pass_frac_of_peak": (p3["algo_bytes"]["sort_pass"] / (p3["kernel_ms"]["sort_pass"] * 1e-3) / 1e9 / peak)
                if p3["kernel_ms"]["sort_pass"] > 0 else None,
                "peak_gbs": peak, "peak_source": peak_src,
            }
        except Exception as e:  # never lose the headline line to the extra measurement
            extra["cfg3_error"] = repr(e)
        try:  # the per-GPU shape of BASELINE config 5 (2^25 uniform cubes) on ONE GPU: the base of the N > 1 arm's large shape
            m5, _ = time_workload(bp, torch, "cfg2", 3, 2, n=1 << 25, with_e2e=False, with_profile=False)
            extra["cfg5_shape_2^25_objects_per_gpu"] = {
                "objects_total": m5["n"], "ms_per_step": m5["ms"], "objects_per_s": m5["n"] / (m5["ms"] * 1e-3),
                "pairs": m5["pairs"], "pairs_per_s": m5["pairs"] / (m5["ms"] * 1e-3), "steps": 3,
                "records": m5["stats"]["n_records"], "sort_passes": m5["stats"]["sort_passes"]}
        except Exception as e:
            extra["cfg5_error"] = repr(e)

    st = m["stats"]
    line = {
        "metric": "objects/sec for extend+sort+scan", "value": value, "unit": "objects/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 keys / u32 ids (f32 quantiser)", "data": "synthetic",
        "config": {"workload": WORKLOADS[wl]["desc"], "objects_per_step": n, "records": st["n_records"],
                   "raw_pairs": st["n_raw_pairs"], "unique_pairs": m["pairs"], "sort_passes": st["sort_passes"],
                   "pair_sort_passes": st["pair_sort_passes"],
                   "l2": "256 MiB buffer written between steps, outside the per-step CUDA-event windows"},
        "pairs_per_s": m["pairs"] / (ms * 1e-3),
        "e2e": {"value": n / (m["e2e_ms"] * 1e-3), "unit": "objects/s", "h2d_bytes_per_step": m["h2d"],
                "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["e2e_ms"],
                "pairs_per_s": m["pairs"] / (m["e2e_ms"] * 1e-3)},
        "gpu_launches": m["launches"],
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks.summary(),
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
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the additional 16M-object measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
