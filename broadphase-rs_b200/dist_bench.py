"""bench.py's N > 1 arm: one process per GPU (torchrun), NCCL over NVLink, weak scaling.

Every rank owns `objects_per_gpu` objects of one global scene (block-distributed by ID, like SURVEY.md
section 8d config 5).  A step is one distributed frame, ONE C-ABI call (bp_dist_frame through dist.DistContext): encode -> sample sort +
all-to-all of the records -> local sort -> halos -> shard-local scan -> all-to-all of the raw pairs ->
sort + dedup.  The result (left sharded on the devices, in rank order) is exactly the reference's
globally sorted, deduplicated pair vector.

Headline shape: 2^25 objects per GPU (BASELINE config 5: 2^28 on 8 GPUs); the latency-bound 2^20-per-GPU shape is
reported beside it.  Before anything is timed the run checks itself (`parity` in the JSON line, rc != 0 on a mismatch):
  * oracle_equal -- one frame of a 2^22-object scene (with a few scene-sized objects, so halos exist) is gathered and
    compared bit for bit with the CPU oracle's par_scan of the whole scene (tests/test_layer.rs:92-124's equality);
  * hash_equal   -- at the timed shape (or, above 2^26 objects in total, on the first 2^26 / N objects of every rank's
    block: one layer holds < 2^30 records and < 2^30 scan work items) a 64-bit order-sensitive hash of the concatenated
    pair list of the N-GPU frame equals that of a 1-GPU frame of the same scene on rank 0;
  * sorted_unique_global -- at the timed shape the pair list is strictly increasing inside and across the ranks' slices."""
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def _i64(c):
    c &= (1 << 64) - 1
    return c - (1 << 64) if c >= (1 << 63) else c


def pair_hash(pairs, first_index):
    """Order-sensitive 64-bit hash of a slice of the pair list: sum over i of mix(pair_i, global index i), mod 2^64.
    pairs: (P, 2) int32 tensor (u32 bit patterns: later ID, earlier ID); first_index: global position of its first pair."""
    if pairs.shape[0] == 0:
        return 0
    p = pairs.to(torch.int64) & 0xFFFFFFFF
    v = (p[:, 0] << 32) | p[:, 1]
    idx = torch.arange(first_index, first_index + p.shape[0], dtype=torch.int64, device=pairs.device)
    x = v ^ (idx * _i64(0x9E3779B97F4A7C15))
    x = x * _i64(0xBF58476D1CE4E5B9)
    x = x ^ (x >> 31)
    x = x * _i64(0x94D049BB133111EB)
    return int(x.sum().item()) & ((1 << 64) - 1)


def _strictly_increasing(pairs):
    if pairs.shape[0] < 2:
        return True
    p = pairs.to(torch.int64) & 0xFFFFFFFF
    v = (p[:, 0] << 32) | p[:, 1]          # < 2^63 while the later ID < 2^31
    return bool((v[1:] > v[:-1]).all().item())


def _scene_slice(bp, n_local, world, rank, seed):
    """This rank's block of a uniform-cube scene of n_local * world objects (recipe of config 2 / 5):
    the cube edge follows the GLOBAL object count, IDs are the global block [rank*n_local, ...)."""
    n_total = n_local * world
    edge_factor = 0.4 * (float(n_total) / float(n_local)) ** (-1.0 / 3.0)
    return bp.scenes.uniform_cubes(n_local, seed + 1000 * rank, id_base=rank * n_local, edge_factor=edge_factor)


def _time_frames(dl, frames, dev_in, host_in, steps, warmup, device, host_path=False, start=0):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev_in[0][0].device)
    stream = torch.cuda.current_stream(device)
    h_pairs = torch.empty((0, 2), dtype=torch.int32).pin_memory() if host_path else None  # pinned landing buffer for the rank's pair slice
    times, pairs_local = [], 0
    for s in range(warmup + steps):
        k = (start + s) % len(frames)
        sc, (d_bounds, d_ids) = frames[k], dev_in[k]
        n_local = sc["bounds"].shape[0]
        flush.fill_(s & 0xFF)
        torch.cuda.synchronize(device)
        dist.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if host_path:
            d_bounds.copy_(host_in[k][0], non_blocking=True)
            d_ids.copy_(host_in[k][1], non_blocking=True)
        pairs = dl.frame(sc["sys_bounds"], d_bounds, d_ids, n_local, None)
        pairs_local = pairs.shape[0]
        if host_path:
            if pairs_local > h_pairs.shape[0]:  # (grows during the warm-up step)
                h_pairs = torch.empty((pairs_local + pairs_local // 4, 2), dtype=torch.int32).pin_memory()
            h_pairs[:pairs_local].copy_(pairs, non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize(device)
        wall = (time.perf_counter() - t0) * 1e3
        if s >= warmup:
            times.append(wall if host_path else e0.elapsed_time(e1))
    t = torch.tensor([sum(times)], dtype=torch.float64, device=flush.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the job is as slow as its slowest rank
    p = torch.tensor([pairs_local], dtype=torch.int64, device=flush.device)
    dist.all_reduce(p, op=dist.ReduceOp.SUM)
    h2d = (frames[0]["bounds"].nbytes + frames[0]["ids"].nbytes) if host_path else 0
    d2h = pairs_local * 8 if host_path else 0
    del flush
    return float(t.item()) / steps, int(p.item()), h2d, d2h


def _device_inputs(frames, device):
    return [(torch.from_numpy(sc["bounds"]).cuda(device), torch.from_numpy(sc["ids"].view(np.int32)).cuda(device)) for sc in frames]


def _host_inputs(frames):
    return [(torch.from_numpy(sc["bounds"]).pin_memory(), torch.from_numpy(sc["ids"].view(np.int32)).pin_memory()) for sc in frames]


def _context(bp, bpd, kind, device, n_local):
    """A bp_dist context sized for n_local objects per rank: at most 8 records per object at min_depth 0, some imbalance
    allowed; the pair buffers grow on demand (collectively, during the warm-up)."""
    return bpd.DistContext(bp, kind, 0, device, record_capacity=int(n_local * 8 * 1.3) + 4096, pair_capacity=6 * n_local + 4096)


def _profile(ops, fn):
    """Runs fn() with per-launch CUDA events on every layer of `ops`; -> summed per-class stats."""
    for l in ops.layers():
        l.set_profiling(True)
        l.reset_stats()
    fn()
    prof = {"kernel_ms": {}, "launches": {}, "algo_bytes": {}}
    for l in ops.layers():
        st = l.stats()
        for k in prof:
            for c, v in st[k].items():
                prof[k][c] = prof[k].get(c, 0) + v
        l.set_profiling(False)
    return prof


# ---- self-checks (untimed) -----------------------------------------------------------------------------------------
def _oracle_check(bp, bpd, scenes, kind, world, rank, device):
    """One frame of a 2^22-object scene against the CPU oracle's par_scan, bit for bit (rank 0 compares)."""
    n_total = 1 << 22
    n_local = n_total // world
    ef = 0.4 * float(world) ** (-1.0 / 3.0)
    sc = scenes.uniform_cubes(n_local, 77 + 1000 * rank, id_base=rank * n_local, edge_factor=ef)
    big = np.random.Generator(np.random.Philox(5 + rank))
    k = 8                                   # a few scene-sized boxes per rank: ancestors of whole shards -> halo copies
    mn = (big.random((k, 3)) * 0.6).astype(np.float32)
    sc["bounds"][:k, :3] = mn
    sc["bounds"][:k, 3:] = mn + np.float32(0.3)
    dl = _context(bp, bpd, kind, device, n_local)
    d_bounds = torch.from_numpy(sc["bounds"]).cuda(device)
    d_ids = torch.from_numpy(sc["ids"].view(np.int32)).cuda(device)
    dl.frame(sc["sys_bounds"], d_bounds, d_ids, n_local, None)          # samples the splitters
    pairs = dl.frame(sc["sys_bounds"], d_bounds, d_ids, n_local, None)  # cached splitters, counts fused with the encode
    halo = torch.tensor([dl.last["halo"]], dtype=torch.int64, device=d_bounds.device)
    dist.all_reduce(halo, op=dist.ReduceOp.SUM)
    got = dl.gather_pairs(pairs)
    # every rank's objects, gathered for the oracle
    b_all = [torch.empty_like(d_bounds) for _ in range(world)]
    i_all = [torch.empty_like(d_ids) for _ in range(world)]
    dist.all_gather(b_all, d_bounds)
    dist.all_gather(i_all, d_ids)
    ok = True
    info = {"objects": n_total, "halo_records": int(halo.item()), "pairs": int(got.shape[0])}
    if rank == 0:
        from oracle import cpu_oracle as co     # the checker: untimed, never on the measured path
        import bench
        co.lib().bpo_set_threads(bench.host_threads())
        o = co.OracleLayer(kind, 4, 0)
        o.extend(sc["sys_bounds"], torch.cat(b_all).cpu().numpy(), torch.cat(i_all).cpu().numpy().view(np.uint32))
        o.par_scan()
        want = o.collisions()
        ok = want.shape == got.shape and bool((want == got.astype(np.uint64)).all())
        info["oracle_pairs"] = int(want.shape[0])
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=d_bounds.device)
    dist.broadcast(flag, 0)
    dl.close()
    return bool(flag.item()), info


def _hash_check(bp, dl, frames, dev_in, world, rank, device, kind):
    """N-GPU frame vs a 1-GPU frame of the same scene on rank 0: order-sensitive hash of the whole pair list, pair count,
    and strict increase across the ranks' slices.  A single layer holds < 2^30 records (<= 5.8 per object with this recipe),
    so above 2^27 objects in total the comparison runs on the first 2^27 / N objects of every rank's block."""
    sc, (d_bounds, d_ids) = frames[0], dev_in[0]
    n_local = sc["bounds"].shape[0]
    n_chk = n_local
    while n_chk * world > (1 << 26):   # (2^26 objects is what one layer is known to hold with every recipe density: < 2^30
        n_chk //= 2                    #  records AND < 2^30 (ancestor, descendant) work items in its scan)
    dev = d_bounds.device
    out = {}
    ok_all = True
    for n_use in sorted({n_local, n_chk}, reverse=True):
        b, i = d_bounds[:n_use], d_ids[:n_use]
        pairs = dl.frame(sc["sys_bounds"], b.contiguous(), i.contiguous(), n_use, None)
        cnt = torch.tensor([pairs.shape[0]], dtype=torch.int64, device=dev)
        cnts = [torch.empty_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        cnts = [int(c.item()) for c in cnts]
        first = sum(cnts[:rank])
        h = torch.tensor([_i64(pair_hash(pairs, first))], dtype=torch.int64, device=dev)
        inc = _strictly_increasing(pairs)
        edge = torch.zeros(4, dtype=torch.int64, device=dev)      # my first and last pair, for the checks across slices
        if pairs.shape[0]:
            edge[:2] = pairs[0].to(torch.int64) & 0xFFFFFFFF
            edge[2:] = pairs[-1].to(torch.int64) & 0xFFFFFFFF
        edges = [torch.empty_like(edge) for _ in range(world)]
        dist.all_gather(edges, edge)
        prev = None
        for r in range(world):
            if cnts[r] == 0:
                continue
            e = [int(x) for x in edges[r].tolist()]
            if prev is not None and not (prev < (e[0], e[1])):
                inc = False
            prev = (e[2], e[3])
        incf = torch.tensor([1 if inc else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(incf, op=dist.ReduceOp.MIN)
        dist.all_reduce(h, op=dist.ReduceOp.SUM)                   # wraps mod 2^64
        total = sum(cnts)
        res = {"objects": n_use * world, "pairs": total, "sorted_unique_global": bool(incf.item()),
               "hash": "%016x" % (int(h.item()) & ((1 << 64) - 1))}
        ok_all = ok_all and res["sorted_unique_global"]
        if n_use == n_chk:   # the same scene on ONE GPU (rank 0), through the plain Layer calls
            b_all = [torch.empty_like(b) for _ in range(world)] if rank == 0 else None
            i_all = [torch.empty_like(i) for _ in range(world)] if rank == 0 else None
            dist.gather(b.contiguous(), b_all, dst=0)
            dist.gather(i.contiguous(), i_all, dst=0)
            same = 1
            if rank == 0:
              try:
                one = bp.LayerBuilder().with_device(device).build(kind, "u32")
                one.set_stream(torch.cuda.current_stream(device).cuda_stream)
                bb, ii = torch.cat(b_all), torch.cat(i_all)
                del b_all, i_all
                one.extend_device(sc["sys_bounds"], bb, ii, bb.shape[0])
                one.par_sort()
                ptr, n1 = one.scan_device(None)
                from broadphase_rs_b200.dist import _view
                p1 = _view(ptr, 2 * n1, torch.int32, dev).view(-1, 2)
                h1 = pair_hash(p1, 0)
                res["single_gpu_pairs"], res["single_gpu_hash"] = int(n1), "%016x" % h1
                same = int(n1 == total and ("%016x" % h1) == res["hash"])
                one.close()
                del bb, ii, p1
                torch.cuda.empty_cache()
              except Exception as e:   # the CHECKER could not run (e.g. out of memory beside the arena): not a mismatch
                res["single_gpu_error"] = repr(e)
                same = 2
            f = torch.tensor([same], dtype=torch.int64, device=dev)
            dist.broadcast(f, 0)
            res["hash_equal"] = None if int(f.item()) == 2 else bool(f.item())
            ok_all = ok_all and res["hash_equal"] is not False
        out["objects_%d" % (n_use * world)] = res
    return ok_all, out


def run(args, bp):
    from broadphase_rs_b200 import dist as bpd
    import bench
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scenes = bench.load_scenes()
    kind = bp.Index64_3D
    wl = "cfg5"

    # ---- self-checks, untimed ----
    parity = {"checked": False}
    if not args.no_parity:
        ok_o, info_o = _oracle_check(bp, bpd, scenes, kind, world, rank, local)
        parity = {"checked": True, "oracle_equal": ok_o, "oracle_scene": info_o}

    inp = bench.make_inputs(scenes, wl, 0, world=world, rank=rank)
    frames = inp["frames"]
    n_local = inp["n"]
    dev_in = _device_inputs(frames, local)
    dl = ops = _context(bp, bpd, kind, local, n_local)   # the whole frame is one C-ABI call: bp_dist_frame
    dl.set_stream(torch.cuda.current_stream(local).cuda_stream)
    if not args.no_parity:
        ok_h, info_h = _hash_check(bp, dl, frames, dev_in, world, rank, local, kind)
        parity.update(hash_equal=ok_h, timed_shape=info_h)

    # ---- the timed shape ----
    with bench.ClockSampler(local) as clocks:
        launches0 = sum(l.stats()["launches_total"] for l in ops.layers())
        ms, pairs, _, _ = _time_frames(dl, frames, dev_in, None, args.steps, args.warmup, local)
        launches = sum(l.stats()["launches_total"] for l in ops.layers()) - launches0
    last = dict(dl.last)
    prof = _profile(ops, lambda: _time_frames(dl, frames, dev_in, None, args.steps, 0, local))
    host_in = _host_inputs(frames)
    e2e_ms, e2e_pairs, h2d, d2h = _time_frames(dl, frames, dev_in, host_in, args.steps, 1, local, host_path=True)
    d2h_t = torch.tensor([d2h], dtype=torch.int64, device="cuda:%d" % local)
    dist.all_reduce(d2h_t, op=dist.ReduceOp.SUM)
    del host_in, dev_in, frames, inp

    extra = {}
    if not args.no_extra:  # the latency-bound small shape: 2^20 objects per GPU (round 1's headline)
        try:
            inp_s = bench.make_inputs(scenes, "cfg2", 0, world=world, rank=rank)
            dev_s = _device_inputs(inp_s["frames"], local)
            dl_s = _context(bp, bpd, kind, local, inp_s["n"])
            dl_s.set_stream(torch.cuda.current_stream(local).cuda_stream)
            ms_s, pairs_s, _, _ = _time_frames(dl_s, inp_s["frames"], dev_s, None, 20, 3, local)
            dl_s.close()
            extra["2^20_objects_per_gpu"] = {
                "objects_per_step": inp_s["n"] * world, "ms_per_step": ms_s, "value": inp_s["n"] * world / (ms_s * 1e-3),
                "unit": "objects/s", "pairs": pairs_s, "pairs_per_s": pairs_s / (ms_s * 1e-3), "steps": 20}
            del dev_s, inp_s
        except Exception as e:
            extra["2^20_error"] = repr(e)

    cpu = None
    if rank == 0:
        from oracle import cpu_oracle as co   # cpu_baseline leg (rank 0 only; the other ranks wait at the barrier below)
        cpu = bench.cpu_baseline(co, scenes, wl, args.cpu_budget, world=world)
    dist.barrier()

    if rank == 0:
        peak, peak_src = bench.hbm_peak()
        n_total = n_local * world
        roof = bench.roofline_block("cfg5_dist", prof, args.steps, peak, peak_src)
        roof["note"] = "rank 0, per-launch CUDA events"
        line = {
            "metric": bench.METRIC, "value": n_total / (ms * 1e-3), "unit": "objects/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": bench.DTYPE, "data": "synthetic",
            "config": bench.workload_config(wl, world),
            "pairs_per_s": pairs / (ms * 1e-3),
            "counts": {"records_owned_rank0": last.get("records_owned"), "halo_records_rank0": last.get("halo"),
                       "raw_pairs_rank0": last.get("raw_pairs"), "unique_pairs": pairs},
            "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "objects/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": int(d2h_t.item()), "ms_per_step": e2e_ms},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity": parity,
            "clocks": clocks.summary(),
        }
        if extra:
            line["other_workloads"] = extra
        bench.emit_line(line)
    dist.barrier()
    ok = (not parity.get("checked")) or (parity.get("oracle_equal") and parity.get("hash_equal"))
    dist.destroy_process_group()
    return 0 if ok else 3
