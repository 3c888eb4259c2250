"""bench.py's N > 1 arm: one process per GPU (torchrun), NCCL over NVLink, weak scaling.

Every rank owns `objects_per_gpu` objects of one global scene (block-distributed by ID, like SURVEY.md
section 8d config 5).  A step is one distributed frame of dist.DistLayer: encode -> sample sort +
all-to-all of the records -> local sort -> halos -> shard-local scan -> all-to-all of the raw pairs ->
sort + dedup.  The result (left sharded on the devices, in rank order) is exactly the reference's
globally sorted, deduplicated pair vector."""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def _scene_slice(bp, n_local, world, rank, seed):
    """This rank's block of a uniform-cube scene of n_local * world objects (recipe of config 2 / 5):
    the cube edge follows the GLOBAL object count, IDs are the global block [rank*n_local, ...)."""
    n_total = n_local * world
    edge_factor = 0.4 * (float(n_total) / float(n_local)) ** (-1.0 / 3.0)
    return bp.scenes.uniform_cubes(n_local, seed + 1000 * rank, id_base=rank * n_local, edge_factor=edge_factor)


def _time_frames(bp, bpd, dl, ops, sc, n_local, steps, warmup, device, host_path=False):
    d_bounds = torch.from_numpy(sc["bounds"]).cuda(device)
    d_ids = torch.from_numpy(sc["ids"].view(np.int32)).cuda(device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=d_bounds.device)
    stream = torch.cuda.current_stream(device)
    if host_path:
        h_bounds = torch.from_numpy(sc["bounds"]).pin_memory()
        h_ids = torch.from_numpy(sc["ids"].view(np.int32)).pin_memory()
        h_pairs = torch.empty((0, 2), dtype=torch.int32).pin_memory()  # pinned landing buffer for the rank's pair slice
    times, pairs_local = [], 0
    for s in range(warmup + steps):
        flush.fill_(s & 0xFF)
        torch.cuda.synchronize(device)
        dist.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if host_path:
            d_bounds.copy_(h_bounds, non_blocking=True)
            d_ids.copy_(h_ids, non_blocking=True)
        pairs = dl.frame(sc["sys_bounds"], d_bounds, d_ids, n_local, None)
        if host_path:
            pairs_local = pairs.shape[0]
            if pairs_local > h_pairs.shape[0]:  # (grows during the warm-up step)
                h_pairs = torch.empty((pairs_local + pairs_local // 4, 2), dtype=torch.int32).pin_memory()
            h_pairs[:pairs_local].copy_(pairs, non_blocking=True)
        else:
            pairs_local = pairs.shape[0]
        e1.record(stream)
        torch.cuda.synchronize(device)
        wall = (time.perf_counter() - t0) * 1e3
        if s >= warmup:
            times.append(wall if host_path else e0.elapsed_time(e1))
    t = torch.tensor([sum(times)], dtype=torch.float64, device=d_bounds.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the job is as slow as its slowest rank
    p = torch.tensor([pairs_local], dtype=torch.int64, device=d_bounds.device)
    dist.all_reduce(p, op=dist.ReduceOp.SUM)
    h2d = (sc["bounds"].nbytes + sc["ids"].nbytes) if host_path else 0
    d2h = pairs_local * 8 if host_path else 0
    del d_bounds, d_ids, flush
    return float(t.item()) / steps, int(p.item()), h2d, d2h


def run(args, bp):
    from broadphase_rs_b200 import dist as bpd
    import bench
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kind = bp.Index64_3D
    n_local = bench.WORKLOADS["cfg2"]["n"]
    ops = bpd.CudaOps(bp, kind, 0, local)
    dl = bpd.DistLayer(ops, kind)
    sc = _scene_slice(bp, n_local, world, rank, 6)

    with bench.ClockSampler(local) as clocks:
        launches0 = sum(l.stats()["launches_total"] for l in ops.layers())
        ms, pairs, _, _ = _time_frames(bp, bpd, dl, ops, sc, n_local, args.steps, args.warmup, local)
        launches = sum(l.stats()["launches_total"] for l in ops.layers()) - launches0
    last = dict(dl.last)
    # per-kernel-class timing on this rank (profiling mode: one CUDA-event pair per launch)
    for l in ops.layers():
        l.set_profiling(True)
        l.reset_stats()
    _time_frames(bp, bpd, dl, ops, sc, n_local, args.steps, 0, local)
    prof = {"kernel_ms": {}, "launches": {}, "algo_bytes": {}}
    for l in ops.layers():
        st = l.stats()
        for k in prof:
            for c, v in st[k].items():
                prof[k][c] = prof[k].get(c, 0) + v
        l.set_profiling(False)
    e2e_ms, e2e_pairs, h2d, d2h = _time_frames(bp, bpd, dl, ops, sc, n_local, args.steps, 1, local, host_path=True)

    extra = {}
    if not args.no_extra:  # the shape of BASELINE config 5: 2^25 objects per GPU (256M at 8 GPUs)
        try:
            n_big = 1 << 25
            sc_big = _scene_slice(bp, n_big, world, rank, 7)
            big_steps = 3
            ms_big, pairs_big, _, _ = _time_frames(bp, bpd, dl, ops, sc_big, n_big, big_steps, 2, local)
            extra["cfg5_shape_2^25_objects_per_gpu"] = {
                "objects_total": n_big * world, "ms_per_step": ms_big, "objects_per_s": n_big * world / (ms_big * 1e-3),
                "pairs": pairs_big, "pairs_per_s": pairs_big / (ms_big * 1e-3), "steps": big_steps,
                "halo_records_rank0": dl.last.get("halo"), "records_owned_rank0": dl.last.get("records_owned")}
            del sc_big
        except Exception as e:
            extra["cfg5_error"] = repr(e)

    if rank == 0:
        peak, peak_src = bench.hbm_peak()
        n_total = n_local * world
        kclass = max((c for c in prof["kernel_ms"] if c != "misc"), key=lambda c: prof["kernel_ms"][c])  # "misc" is a grab-bag of small helpers, not one kernel
        k_ms, k_launch, k_bytes = prof["kernel_ms"][kclass], prof["launches"][kclass], prof["algo_bytes"][kclass]
        achieved = (k_bytes / (k_ms * 1e-3) / 1e9) if k_ms > 0 else 0.0
        line = {
            "metric": "objects/sec for extend+sort+scan", "value": n_total / (ms * 1e-3), "unit": "objects/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 keys / u32 ids (f32 quantiser)", "data": "synthetic",
            "config": {"workload": "%d x 2^20 uniform-size AABBs (edge 0.4*N_total^-1/3), Index64_3D, range-sharded by Morton "
                                   "prefix: sample sort + all-to-all, ancestor halos, global pair dedup" % world,
                       "objects_per_step": n_total, "objects_per_gpu": n_local, "parallelism": "morton-range-shard x%d" % world,
                       "records_owned_rank0": last.get("records_owned"), "halo_records_rank0": last.get("halo"),
                       "unique_pairs": pairs,
                       "l2": "256 MiB buffer written between steps, outside the per-step CUDA-event windows"},
            "pairs_per_s": pairs / (ms * 1e-3),
            "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "objects/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_ms},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": kclass, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": bench.ncu_traffic(kclass), "peak_source": peak_src,
                         "per_class_ms_per_step": {c: v / args.steps for c, v in prof["kernel_ms"].items()},
                         "note": "rank 0, per-launch CUDA events"},
            "cpu_baseline": None,
            "clocks": clocks.summary(),
        }
        if extra:
            line["other_workloads"] = extra
        bench.emit_line(line)
    dist.barrier()
    dist.destroy_process_group()
    return 0
