"""world_size > 1 tests of the multi-GPU choreography (tests/dist_protocol.DistLayer, the executable model of bp_dist_frame's protocol) on CPU: gloo backend, the
numpy/oracle test double for the shard-local operations.  The concatenation of the ranks' results
must equal the single-process oracle scan bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from tests import dist_cpu_ops as dco

CASES = ["uniform3d", "big_objects3d", "multibounds2d", "multibounds3d", "skewed3d"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,empty_rank", [(2, -1), (3, -1), (2, 1)])
def test_distributed_frame_equals_single_process_oracle(tmp_path, world, empty_rank):
    mp.spawn(dco.worker, args=(world, _free_port(), CASES, empty_rank, str(tmp_path)), nprocs=world, join=True)
    for case in CASES:
        got = np.load(os.path.join(str(tmp_path), "%s.npy" % case))
        want = dco.reference_pairs(case)
        assert got.shape == want.shape, (case, got.shape, want.shape)
        assert (got == want).all(), case
    for frame in range(2):  # static layer sharded once + dynamic layer merged per frame (config 4 at N > 1)
        got = np.load(os.path.join(str(tmp_path), "static_dynamic_%d.npy" % frame))
        want = dco.reference_static_dynamic(frame)
        assert got.shape == want.shape and (got == want).all(), ("static+dynamic", frame)
    halos = np.load(os.path.join(str(tmp_path), "big_objects3d.halos.npy"))
    if empty_rank < 0:
        assert halos[0] == 0 and halos[1:].sum() > 0  # scene-sized objects must have produced halo records


@pytest.mark.parametrize("world,empty_rank", [(2, -1), (3, 1)])
def test_product_protocol_plans_hold_on_the_received_records(tmp_path, world, empty_rank):
    """DistLayer.frame through the protocol the CUDA path speaks (tag words in the count matrix, sort planned from them,
    counts fused with the encode once the splitters are cached), on the CPU double: the double checks every plan against the
    records that actually arrived (masks cover them; "IDs ascending" is true of the buffer), and the pairs are the oracle's."""
    import json
    mp.spawn(dco.worker, args=(world, _free_port(), CASES, empty_rank, str(tmp_path), True), nprocs=world, join=True)
    for case in CASES:
        got = np.load(os.path.join(str(tmp_path), "%s.npy" % case))
        want = dco.reference_pairs(case)
        assert got.shape == want.shape and (got == want).all(), case
    for frame in range(2):
        got = np.load(os.path.join(str(tmp_path), "static_dynamic_%d.npy" % frame))
        want = dco.reference_static_dynamic(frame)
        assert got.shape == want.shape and (got == want).all(), ("static+dynamic", frame)
    # a scene change under cached splitters: frames 0-1 scene A, frames 2-4 scene B; frame 2 runs on stale splitters
    want_a, want_b = dco.reference_pairs_unfiltered("uniform3d"), dco.reference_pairs_unfiltered("skewed3d")
    for f in range(5):
        got = np.load(os.path.join(str(tmp_path), "rebalance_%d.npy" % f))
        want = want_a if f < 2 else want_b
        assert got.shape == want.shape and (got == want).all(), ("rebalance frame", f)
    # fused: frames 1, 2 and 4 -- frame 3 had to sample new splitters (so the imbalance was noticed and acted on)
    assert int(np.load(os.path.join(str(tmp_path), "rebalance_fused.npy"))[0]) == 3
    stats = json.load(open(os.path.join(str(tmp_path), "product_stats.json")))
    plans = [p for rank in stats for fused, ps in rank for p in ps if p[1] > 0]
    assert all(fused >= 1 for rank in stats for fused, _ in rank)          # cached splitters -> counts came with the encode
    assert any(asc for asc, _ in plans) and any(not asc for asc, _ in plans)  # both kinds of plan were exercised


def test_ancestor_keys_and_splitters(bp):
    from tests import dist_protocol as bpd
    from oracle import cpu_oracle as co
    key = co.make_index(2, 5, [0x12345678 & 0xF8000000, 0x9abcdef0 & 0xF8000000, 0x0fedcba9 & 0xF8000000])  # origin truncated to depth 5
    anc = bpd.ancestor_keys(2, key)
    assert len(anc) == 6 and anc[0] == 0 and anc[-1] == key and anc == sorted(anc)
    for d, a in enumerate(anc):
        assert a & 31 == d and co.overlaps(2, a, key)
    s = bpd.choose_splitters(np.arange(1000, dtype=np.uint64), 4)
    assert list(s) == [256, 496, 752]  # the roundest values within 1000 / 4 / 32 = 7 sample ranks of the quantiles 250, 500, 750
    assert list(bpd.choose_splitters(np.arange(100, dtype=np.uint64), 4)) == [25, 50, 75]  # too small a sample to move anything
    assert list(bpd.choose_splitters(np.zeros(0, dtype=np.uint64), 4)) == [2**64 - 1] * 3
    # run_upper_key: the cell of `key` at depth 5 spans the low 3*(19-5) origin bits + the depth field
    assert bpd.run_upper_key(2, key) == key | ((1 << (5 + 3 * 14)) - 1)
    m_own = np.array([[5, 1], [2, 7]])
    m_halo = np.array([[0, 3], [0, 0]])
    assert bpd.chunk_offsets(m_own, m_halo, 0) == ([0, 0], [5, 1])
    assert bpd.chunk_offsets(m_own, m_halo, 1) == ([5, 4], [7, 11])


def test_round_splitters_and_the_bits_a_shard_shares(bp):
    """Splitters snapped to round values stay ascending and within the slack of their quantiles; shard_fixed_bits is true
    of every key that the shard rule (#splitters <= key) sends to the shard, for keys AND for 32-bit later IDs."""
    from tests import dist_protocol as bpd
    rng = np.random.Generator(np.random.Philox(5))
    full = (1 << 64) - 1
    for trial in range(60):
        parts = int(rng.integers(2, 17))
        bits = int(rng.integers(4, 65))
        m = int(rng.integers(1, 40000))
        sample = rng.integers(0, 1 << bits, size=m, dtype=np.uint64, endpoint=False) if bits < 64 else rng.integers(0, full, size=m, dtype=np.uint64, endpoint=True)
        if trial % 5 == 0:  # heavy duplicates
            sample = sample >> np.uint64(max(0, bits - 3)) << np.uint64(max(0, bits - 3))
        top = (1 << int(np.bitwise_or.reduce(sample)).bit_length()) - 1  # what the product knows: the sources' OR
        spl = bpd.choose_splitters(sample, parts)
        assert spl.shape == (parts - 1,) and (np.diff(spl.astype(object)) >= 0).all()
        srt = np.sort(sample)
        slack = m // parts // 32
        for i, v in enumerate(spl, start=1):
            t = min(m - 1, i * m // parts)
            assert int(srt[max(0, t - slack)]) <= int(v) <= int(srt[min(m - 1, t + slack)])
        shard = np.searchsorted(spl, sample, side="right")
        for d in range(parts):
            fixed, value = bpd.shard_fixed_bits(spl, d, top)
            mine = sample[shard == d]
            assert value & ~fixed == 0
            assert ((mine & np.uint64(fixed)) == np.uint64(value)).all(), (trial, d)
    # uniform keys, power-of-two shards, a sample large enough that the octant boundaries fall inside the windows: every
    # shard shares its top log2(parts) bits (8 shards of the 62-bit keys: 3 bits less to sort on, each)
    sample = rng.integers(0, 1 << 62, size=1 << 19, dtype=np.uint64)
    for parts in (2, 4, 8, 16):
        spl = bpd.choose_splitters(sample, parts)
        assert [int(v) for v in spl] == [i << (62 - parts.bit_length() + 1) for i in range(1, parts)]
        for d in range(parts):
            fixed, value = bpd.shard_fixed_bits(spl, d, (1 << 62) - 1)
            assert fixed == full & ~((1 << (62 - parts.bit_length() + 1)) - 1) and value == d << (62 - parts.bit_length() + 1)
    # degenerate splitters: empty shards share nothing, the last shard's upper end is `top`
    assert bpd.shard_fixed_bits(np.array([8, 8, 16], dtype=np.uint64), 1, 31) == (0, 0)
    assert bpd.shard_fixed_bits(np.array([8, 8, 16], dtype=np.uint64), 2, 31) == (full & ~7, 8)
    assert bpd.shard_fixed_bits(np.array([8, 8, 16], dtype=np.uint64), 3, 31) == (full & ~15, 16)
    assert bpd.shard_fixed_bits(np.array([8, 8, 16], dtype=np.uint64), 0, 31) == (full & ~7, 0)
    assert bpd.shard_fixed_bits(np.array([full] * 3, dtype=np.uint64), 0) == (0, 0)  # splitters of an empty sample: everything in shard 0


def test_product_splitter_planner_equals_the_model(bp):
    """The C++ planner of bp_dist_frame (bp_dist_plan_splitters / bp_dist_plan_shard_bits: host code, no device needed)
    against tests/dist_protocol.py on random samples -- duplicates, tiny samples, every shard count, 64- and 32-bit tops."""
    from tests import dist_protocol as bpd
    rng = np.random.Generator(np.random.Philox(23))
    full = (1 << 64) - 1
    for trial in range(200):
        parts = int(rng.integers(1, 17))
        bits = int(rng.integers(1, 65))
        m = int(rng.integers(0, 5000)) if trial % 3 else int(rng.integers(0, 40))
        hi = full if bits == 64 else (1 << bits) - 1
        sample = rng.integers(0, hi, size=m, dtype=np.uint64, endpoint=True)
        if trial % 4 == 0 and m:
            sample = sample[rng.integers(0, max(1, m // 50), size=m)]  # few distinct values
        want = bpd.choose_splitters(sample, parts)
        got = bp.plan_dist_splitters(sample, parts)
        assert got.shape == want.shape and (got == want).all(), (trial, parts, m)
        top = [full, 0xFFFFFFFF, (1 << max(1, int(np.bitwise_or.reduce(sample)).bit_length())) - 1 if m else 0][trial % 3]
        for d in range(parts):
            assert bp.plan_dist_shard_bits(want, d, top) == bpd.shard_fixed_bits(want, d, top), (trial, d, [hex(int(v)) for v in want], hex(top))
    with pytest.raises(bp.BpError):
        bp.plan_dist_splitters(np.zeros(4, dtype=np.uint64), 17)
    with pytest.raises(bp.BpError):
        bp.plan_dist_shard_bits(np.zeros(3, dtype=np.uint64), 4)


def test_sort_plan_from_tag_words(bp):
    """The receivers' sort plan from the senders' tag words: masks OR / AND over the sources, IDs ascending only if every
    source's are, the ID ranges follow each other in rank order and no halo copies arrived."""
    from tests.dist_protocol import sort_plan
    full = (1 << 64) - 1
    empty = [0, 0, full, full, full, 0, 1]
    a = [0xFF | (1 << 63), 0xF0F0, 0x1010, 0x01, 0, 99, 1]
    b = [0x1FF, 0x0F0F, 0x0101, 0x100, 100, 250, 1]
    assert sort_plan([a, b], 0x1FF, 0) == (0xFFFF, 0, 0x1FF, 0, True)
    assert sort_plan([a, empty, b], 0x1FF, 0)[4] is True
    assert sort_plan([b, a], 0x1FF, 0)[4] is False            # ranges out of rank order
    assert sort_plan([a, b], 0x1FF, 3)[4] is False            # halo copies are unordered
    assert sort_plan([a, b[:6] + [0]], 0x1FF, 0)[4] is False  # a source whose own IDs do not ascend
    assert sort_plan([a, [0x1FF, 0, full, full, 99, 250, 1]], 0x1FF, 0)[4] is True   # equal IDs may meet at the seam
    assert sort_plan([empty, empty], 0, 0) == (0, full, 0, full, True)
    # the receiving shard's keys share the bits its two splitters share -- unless halo copies (below the lower splitter) came
    spl = np.array([0x3000, 0x4000, 0xC000], dtype=np.uint64)
    assert sort_plan([a, b], 0x1FF, 0, (spl, 1))[:2] == (0x3FFF, 0x3000)   # [0x3000, 0x4000): the top 4 of 16 bits are 0011
    assert sort_plan([a, b], 0x1FF, 2, (spl, 1))[:2] == (0xFFFF, 0)
    assert sort_plan([a, b], 0x1FF, 0, (spl, 3))[:2] == (0xFFFF, 0xC000)   # the last shard ends at the sources' largest possible key
    assert sort_plan([a, b], 0x1FF, 0, (spl, 0))[:2] == (0x3FFF, 0)


def test_scatter_destinations_equal_the_per_destination_sums(bp):
    """The array arithmetic that runs between the count matrix and the scatter launch (GPU idle) against the plain
    per-destination sums it replaced; addresses stay exact 64-bit integers."""
    from tests import dist_protocol as bpd
    rng = np.random.Generator(np.random.Philox(77))
    for g in (1, 2, 3, 8, 16):
        kp = (np.uint64(0x7F00_0000_0000) + np.arange(g, dtype=np.uint64) * np.uint64(1 << 36))
        ip = kp + np.uint64(1 << 35)
        for with_halo in (False, True):
            m_own = rng.integers(0, 1 << 28, size=(g, g)).astype(np.int64)
            m_halo = (rng.integers(0, 50, size=(g, g)) * np.triu(np.ones((g, g), dtype=np.int64), 1)).astype(np.int64)
            if not with_halo:
                m_halo[:] = 0
            for me in range(g):
                dk, di, hk, hi = bpd.scatter_destinations(kp, ip, m_own, m_halo, me)
                own = [sum(int(m_own[s, d]) + int(m_halo[s, d]) for s in range(me)) for d in range(g)]
                assert dk.dtype == np.uint64 and di.dtype == np.uint64
                assert dk.tolist() == [int(kp[d]) + 8 * own[d] for d in range(g)]
                assert di.tolist() == [int(ip[d]) + 4 * own[d] for d in range(g)]
                assert bpd.chunk_offsets(m_own, m_halo, me) == (own, [own[d] + int(m_own[me, d]) for d in range(g)])
                if m_halo[me].any():
                    assert hk.dtype == np.uint64 and hi.dtype == np.uint64
                    assert hk.tolist() == [int(kp[d]) + 8 * (own[d] + int(m_own[me, d])) for d in range(g)]
                    assert hi.tolist() == [int(ip[d]) + 4 * (own[d] + int(m_own[me, d])) for d in range(g)]
                else:
                    assert hk is None and hi is None
