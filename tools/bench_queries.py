"""Throughput of the batched queries (Layer::test_box / test_ray) on the BASELINE config-2 scene (2^20 uniform
cubes, Index64_3D): queries/s on the GPU through the C ABI (device-resident geometry, results left on the
device; CUDA events) next to the CPU oracle's single-query loop on a bounded sample.  One JSON line.
Usage: python tools/bench_queries.py [n_objects_log2=20] [n_queries_log2=20]"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import _loadpkg
from oracle import cpu_oracle as co

bp = _loadpkg.load()
from importlib import import_module

lib = import_module(bp.__name__ + "._lib").lib
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
nq = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
sc = bp.scenes.uniform_cubes(n, 2)
sysb = np.ascontiguousarray(sc["sys_bounds"], dtype=np.float32)
rng = np.random.Generator(np.random.Philox(11))
edge = np.float32(2.0 * n ** (-1.0 / 3.0))            # a box a few objects wide
mn = (rng.random((nq, 3)) * (1.0 - edge)).astype(np.float32)
boxes = np.concatenate([mn, mn + edge], axis=1).astype(np.float32)
org = rng.random((nq, 3)).astype(np.float32)
d = rng.normal(size=(nq, 3)).astype(np.float32)
rays = np.concatenate([org, d, np.zeros((nq, 1), np.float32), np.full((nq, 1), 0.05, np.float32)], axis=1).astype(np.float32)

stream = torch.cuda.current_stream()
L = bp.Layer(bp.Index64_3D, "u32")
L.set_stream(stream.cuda_stream)
L.extend(sysb, sc["bounds"], sc["ids"])
L.par_sort()
out = {"objects": n, "queries": nq, "records": len(L)}


def gpu(fn, params, steps=5):
    dq = torch.from_numpy(params).cuda()
    pairs, offs, cnt = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
    ts = []
    for it in range(steps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = fn(L._h, sysb.ctypes.data, dq.data_ptr(), params.shape[0], -1, 1, ctypes.byref(pairs), ctypes.byref(offs), ctypes.byref(cnt))
        assert st == 0, st
        e1.record(stream)
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts), cnt.value


def cpu(kind, params, sample=2000):
    o = co.OracleLayer(co.INDEX64_3D, 4, 0)
    o.extend(sysb, sc["bounds"], sc["ids"])
    o.sort()
    t0 = time.perf_counter()
    tot = 0
    for q in params[:sample]:
        if kind == "box":
            tot += o.test_box(sysb, q).shape[0]
        else:
            tot += o.test_ray(sysb, q[:3], q[3:6], q[6], q[7]).shape[0]
    dt = time.perf_counter() - t0
    return sample / dt, tot / sample


for kind, fn, params in (("box", lib().bp_layer_test_box_batch, boxes), ("ray", lib().bp_layer_test_ray_batch, rays)):
    ms, cnt = gpu(fn, params)
    qps_cpu, avg = cpu(kind, params)
    out[kind] = {"gpu_ms": ms, "gpu_queries_per_s": nq / (ms * 1e-3), "results": cnt, "results_per_query": cnt / nq,
                 "cpu_oracle_queries_per_s_1_thread": qps_cpu, "cpu_sample_results_per_query": avg}
print(json.dumps(out))
