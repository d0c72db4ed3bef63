"""GPU: vnode (CRC32) / dispatcher-rewrite / stable-partition kernels vs the oracle and numpy."""
import zlib

import numpy as np
import pytest

from risingwave_b200 import abi
from risingwave_b200.stream_chunk import Column, StreamChunk

from helpers import rand_chunk

pytestmark = pytest.mark.gpu


def test_vnode_matches_oracle_all_types(cuda, oracle):
    rng = np.random.default_rng(0)
    types = [abi.T_INT16, abi.T_INT32, abi.T_INT64, abi.T_FLOAT32, abi.T_FLOAT64, abi.T_BOOL, abi.T_TIMESTAMPTZ, abi.T_DATE]
    n = 5000
    ch = rand_chunk(rng, n, types, null_frac=0.1, vis_frac=0.9)
    ch.columns[5] = Column(abi.T_BOOL, rng.integers(0, 2, n).astype(np.uint8), ch.columns[5].valid)
    f = ch.columns[4].data
    f[:5] = [0.0, -0.0, np.nan, np.inf, -np.inf]
    for keys in ([0], [1], [2], [3], [4], [5], [2, 1], [0, 1, 2, 3, 4, 5, 6, 7]):
        for vc in (256, 1024, 4096):
            got = cuda.vnode_compute(ch, keys, vc)
            want = oracle.vnode_compute(ch, keys, vc)
            assert np.array_equal(got, want), (keys, vc)


def test_vnode_int_keys_match_reference_formula(cuda):
    """test_hash_dispatcher (dispatch.rs:1593-1606): crc32 over the LE bytes of the i32 key columns % 256."""
    rng = np.random.default_rng(1)
    n = 4096
    a = rng.integers(-2**31, 2**31, n).astype(np.int32)
    c = rng.integers(0, 10, n).astype(np.int32)
    ch = StreamChunk(np.full(n, 1, np.uint8), [Column(abi.T_INT32, a), Column(abi.T_INT32, c)])
    got = cuda.vnode_compute(ch, [0, 1], 256)
    want = [zlib.crc32(a[i].tobytes() + c[i].tobytes()) % 256 for i in range(n)]
    assert got.tolist() == want


def test_vnode_serial_row_id(cuda, oracle):
    rng = np.random.default_rng(2)
    n = 1000
    ids = rng.integers(0, 2**62, n).astype(np.int64)
    valid = rng.random(n) > 0.1
    ch = StreamChunk(np.full(n, 1, np.uint8), [Column(abi.T_SERIAL, ids, valid), Column(abi.T_INT64, ids // 7)])
    for vc in (256, 2048):
        assert np.array_equal(cuda.vnode_compute(ch, [0], vc), oracle.vnode_compute(ch, [0], vc))


def test_dispatch_rewrite_ops(cuda, oracle):
    ch = StreamChunk.from_pretty(" I I\n U- 1 10\n U+ 1 11\n U- 2 20\n U+ 3 20\n + 4 0\n U- 5 1\n + 9 9 D\n U+ . 1\n - 7 7")
    assert cuda.dispatch_rewrite_ops(ch, [0]).tolist() == oracle.dispatch_rewrite_ops(ch, [0]).tolist()
    bad = StreamChunk.from_pretty(" I I\n U+ 1 10")
    with pytest.raises(abi.RwError):
        cuda.dispatch_rewrite_ops(bad, [0])


def test_stable_partition_device(cuda, oracle):
    import torch
    from risingwave_b200 import device, exchange
    rng = np.random.default_rng(3)
    for n, world in ((1, 2), (2047, 2), (2048, 8), (100_000, 8), (300_001, 4)):
        key = rng.integers(0, 5000, n).astype(np.int64)
        pay = np.arange(n, dtype=np.int64)
        small = rng.integers(0, 100, n).astype(np.int32)
        ops = rng.integers(1, 5, n).astype(np.uint8)
        ops[rng.random(n) < 0.05] = 0  # rows folded to "invisible" are dropped
        chunk = device.DeviceChunk(torch.from_numpy(ops).cuda(),
                                   [torch.from_numpy(key).cuda(), torch.from_numpy(pay).cuda(), torch.from_numpy(small).cuda()],
                                   [abi.T_INT64, abi.T_INT64, abi.T_INT32])
        v2d = exchange.vnode_to_dest_table(world).cuda()
        o_ops, o_cols, counts, offsets = device.shuffle_partition(chunk, [0], v2d, world)
        torch.cuda.synchronize()
        host = StreamChunk(np.where(ops == 0, 1, ops).astype(np.uint8), [Column(abi.T_INT64, key)])
        vnode = oracle.vnode_compute(host, [0], 256).astype(np.int64)
        dest = (vnode * world // 256)
        keep = ops != 0
        order = np.argsort(dest[keep], kind="stable")
        idx = np.nonzero(keep)[0][order]
        cnt = np.bincount(dest[keep], minlength=world)
        assert counts.cpu().numpy().tolist() == cnt.tolist()
        assert offsets.cpu().numpy().tolist() == np.concatenate([[0], np.cumsum(cnt)[:-1]]).tolist()
        m = int(cnt.sum())
        assert np.array_equal(o_ops.cpu().numpy()[:m], ops[idx])
        assert np.array_equal(o_cols[0].cpu().numpy()[:m], key[idx])
        assert np.array_equal(o_cols[1].cpu().numpy()[:m], pay[idx])   # stable: row order kept per destination
        assert np.array_equal(o_cols[2].cpu().numpy()[:m], small[idx])


# ------------------------------------------------------------------------------------------ P2P exchange (round 2)
def _expected_partition(oracle, ops, cols, world):
    """numpy restatement of HashDataDispatcher::dispatch_data's routing (dispatch.rs:961-1053) for the contiguous
    vnode -> destination mapping: per destination, the visible rows in input order."""
    host = StreamChunk(np.where(ops == 0, 1, ops).astype(np.uint8), [Column(abi.T_INT64, cols[0])])
    dest = oracle.vnode_compute(host, [0], 256).astype(np.int64) * world // 256
    keep = ops != 0
    return [np.nonzero(keep & (dest == d))[0] for d in range(world)]


def test_p2p_partition_and_unpack_virtual_ranks(cuda, oracle):
    """part_scatter_p2p_kernel / p2p_publish_counts_kernel / p2p_unpack_kernel -- the default N>1 data path of bench.py --
    with W virtual ranks on ONE device: every rank's partition kernel stores into the W receive buffers, every receiver
    unpacks its W regions; the result must be the numpy stable partition, sources concatenated in rank order."""
    import ctypes as C
    import torch
    from risingwave_b200 import device, exchange
    rng = np.random.default_rng(8)
    types = [abi.T_INT64, abi.T_INT64, abi.T_INT32]
    for world, n in ((2, 5000), (4, 70001), (8, 3)):
        cap = n  # worst case: every row of a source goes to one destination
        region = device.p2p_region_bytes(types, cap)
        bufs = [torch.zeros(world * region, dtype=torch.uint8, device="cuda") for _ in range(world)]
        peers = [b.data_ptr() for b in bufs]
        v2d = exchange.vnode_to_dest_table(world).cuda()
        src = []
        for r in range(world):
            key = rng.integers(0, 100000, n).astype(np.int64)
            pay = (np.arange(n, dtype=np.int64) + r * 10 ** 9)
            small = rng.integers(0, 100, n).astype(np.int32)
            ops = rng.integers(1, 5, n).astype(np.uint8)
            ops[rng.random(n) < 0.05] = 0
            src.append((ops, [key, pay, small]))
            chunk = device.DeviceChunk(torch.from_numpy(ops).cuda(), [torch.from_numpy(c).cuda() for c in (key, pay, small)], types)
            counts = torch.zeros(world, dtype=torch.int64, device="cuda")
            overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
            device.shuffle_partition_p2p(chunk, [0], v2d, world, r, peers, cap, counts, overflow)
            torch.cuda.synchronize()
            assert int(overflow.item()) == 0
            want_idx = _expected_partition(oracle, ops, [key], world)
            assert counts.cpu().numpy().tolist() == [len(ix) for ix in want_idx]
        for d in range(world):
            out_ops = torch.empty(world * cap, dtype=torch.uint8, device="cuda")
            out_cols = [torch.empty(world * cap, dtype=device.TORCH_DTYPE[t], device="cuda") for t in types]
            total = torch.zeros(1, dtype=torch.int64, device="cuda")
            device.shuffle_unpack(peers[d], world, types, cap, out_ops, out_cols, total)
            torch.cuda.synchronize()
            w_ops, w_cols = [], [[] for _ in types]
            for r in range(world):
                ops, cols = src[r]
                ix = _expected_partition(oracle, ops, [cols[0]], world)[d]
                w_ops.append(ops[ix])
                for k in range(len(types)):
                    w_cols[k].append(cols[k][ix])
            w_ops = np.concatenate(w_ops)
            m = int(total.item())
            assert m == len(w_ops)
            assert np.array_equal(out_ops.cpu().numpy()[:m], w_ops)
            for k in range(len(types)):
                assert np.array_equal(out_cols[k].cpu().numpy()[:m], np.concatenate(w_cols[k]))


def test_p2p_region_overflow_is_reported_not_silent(cuda):
    """a (source, destination) pair larger than its region: the overflow flag is raised and the receiver's total is -1
    (the caller sizes regions for the worst case, exchange.P2PShufflePlan, so this never happens on the bench path)."""
    import torch
    from risingwave_b200 import device, exchange
    types = [abi.T_INT64]
    n, cap, world = 4096, 64, 2
    region = device.p2p_region_bytes(types, cap)
    bufs = [torch.zeros(world * region, dtype=torch.uint8, device="cuda") for _ in range(world)]
    chunk = device.DeviceChunk(torch.ones(n, dtype=torch.uint8, device="cuda"), [torch.arange(n, dtype=torch.int64, device="cuda")], types)
    counts = torch.zeros(world, dtype=torch.int64, device="cuda")
    overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
    device.shuffle_partition_p2p(chunk, [0], exchange.vnode_to_dest_table(world).cuda(), world, 0, [b.data_ptr() for b in bufs], cap, counts, overflow)
    torch.cuda.synchronize()
    assert int(overflow.item()) == 1
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    device.shuffle_unpack(bufs[0].data_ptr(), 1, types, cap, torch.empty(cap, dtype=torch.uint8, device="cuda"),
                          [torch.empty(cap, dtype=torch.int64, device="cuda")], total)
    torch.cuda.synchronize()
    assert int(total.item()) == -1


def test_p2p_exchange_self_peer_feeds_counted_join(cuda, oracle):
    """world = 1 through the ONE-CALL exchange (rwgpu_shuffle_exchange_p2p_device: hist, scan, scatter, publish, device
    barrier, unpack, count) into rwgpu_join_push_device_counted, row count read on the device -- the chain bench.py times
    at N>1 -- against the oracle fed the same rows."""
    import ctypes as C
    import torch
    from risingwave_b200 import device, exchange
    from risingwave_b200.executor import HashJoinExecutor, JoinParams, MockSource
    from risingwave_b200.stream_chunk import net_multiset
    rng = np.random.default_rng(12)
    types = [abi.T_INT64] * 4
    nb, n, cap = 3000, 20000, 20000
    exs = []
    for be in (cuda, oracle):
        _, sl = MockSource.channel()
        _, sr = MockSource.channel()
        exs.append(HashJoinExecutor(be, abi.JOIN_INNER, sl.into_executor(types, [1]), sr.into_executor(types, [0]),
                                    JoinParams([0], [1]), JoinParams([0], []), [False], capacity_hint=nb))
    auct = [np.arange(nb, dtype=np.int64)] + [rng.integers(0, 1000, nb).astype(np.int64) for _ in range(3)]
    a_ops = np.full(nb, abi.OP_INSERT, np.uint8)
    for ex in exs:
        assert ex.eq_join_oneside(1, StreamChunk(a_ops, [Column(abi.T_INT64, c) for c in auct])) == []
    region = device.p2p_region_bytes(types, cap)
    recv = torch.zeros(region, dtype=torch.uint8, device="cuda")
    flags = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    out_ops = torch.empty(cap, dtype=torch.uint8, device="cuda")
    out_cols = [torch.empty(cap, dtype=torch.int64, device="cuda") for _ in types]
    counts = torch.zeros(1, dtype=torch.int64, device="cuda")
    overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
    total_host = torch.zeros(1, dtype=torch.int64).pin_memory()
    call = device.P2PExchangeCall([0], exchange.vnode_to_dest_table(1).cuda(), 1, 0, [recv.data_ptr()], [flags.data_ptr()], cap,
                                  recv.data_ptr(), out_ops, out_cols, counts, overflow, total_host)
    stream = torch.cuda.Stream()
    for epoch in (1, 2, 3):
        m = n - 1000 * epoch
        bid = [rng.integers(0, nb + 200, m).astype(np.int64), (np.arange(m) + 10 ** 6 * epoch).astype(np.int64),
               rng.integers(0, 1 << 30, m).astype(np.int64), rng.integers(0, 1 << 30, m).astype(np.int64)]
        ops = np.full(m, abi.OP_INSERT, np.uint8)
        ops[rng.integers(0, m, 40)] = 0
        chunk = device.DeviceChunk(torch.from_numpy(ops).cuda(), [torch.from_numpy(c).cuda() for c in bid], types)
        with torch.cuda.stream(stream):
            call(chunk, epoch, stream)
            count_ptr = flags.data_ptr() + 512 + 8 * (epoch & 1)
            view = device.join_push_device(exs[0], abi.SIDE_LEFT, device.DeviceChunk(out_ops, out_cols, types), stream, n_rows_dev=count_ptr)
            stream.synchronize()
        assert int(total_host.item()) == int((ops != 0).sum())
        vis = view.visible()
        got_ops = view.ops()
        got_cols = [view.column(k) for k in range(view.n_cols)]
        if vis is not None:
            got_ops, got_cols = got_ops[vis], [c[vis] for c in got_cols]
        from collections import Counter
        got = Counter()
        go, gc = got_ops.cpu().numpy(), [c.cpu().numpy() for c in got_cols]
        for i in range(len(go)):
            got[tuple(int(c[i]) for c in gc)] += 1 if go[i] in (abi.OP_INSERT, abi.OP_UPDATE_INSERT) else -1
        keep = ops != 0
        want = net_multiset(exs[1].eq_join_oneside(0, StreamChunk(ops[keep], [Column(abi.T_INT64, c[keep]) for c in bid])))
        assert {k: v for k, v in got.items() if v} == dict(want), f"epoch {epoch}"


def _two_gpu_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        from risingwave_b200 import abi as _abi, device, exchange
        types = [_abi.T_INT64] * 2
        n = 50000
        rng = np.random.default_rng(100 + rank)
        res = {}
        for name in ("flat", "p2p", "nccl"):
            plan = (exchange.FlatShufflePlan(world, rank, [0], types, batch_rows=n) if name == "flat" else
                    exchange.P2PShufflePlan(world, rank, [0], types, batch_rows=n) if name == "p2p" else exchange.ShufflePlan(world, rank, [0], types))
            out = []
            for step in range(3):
                key = rng.integers(0, 1 << 40, n).astype(np.int64)
                if step == 2:
                    key[:] = 7  # extreme skew: every row of every rank goes to ONE destination
                pay = np.arange(n, dtype=np.int64) + rank * 10 ** 9 + step * 10 ** 6
                ch = device.DeviceChunk(torch.ones(n, dtype=torch.uint8, device="cuda"), [torch.from_numpy(key).cuda(), torch.from_numpy(pay).cuda()], types)
                ops, cols = plan.exchange(ch, torch.cuda.current_stream())
                torch.cuda.synchronize()
                out.append((key, pay, ops.cpu().numpy().copy(), [c.cpu().numpy().copy() for c in cols]))
            res[name] = out
        q.put((rank, res))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_p2p_exchange_two_gpus_matches_numpy_and_nccl(cuda, oracle):
    """two real ranks: the one-kernel exchange (FlatShufflePlan, bench.py's default at N>1), the region exchange
    (P2PShufflePlan) and the NCCL all-to-all-v path (ShufflePlan) all deliver
    exactly the rows numpy's stable partition assigns to each rank, sources in rank order -- including a batch whose
    rows ALL go to one destination (regions are sized for that)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29617
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for name in ("flat", "p2p", "nccl"):
        for step in range(3):
            for d in range(world):
                w_pay = []
                for r in range(world):
                    key, pay, _, _ = got[r][name][step]
                    ix = _expected_partition(oracle, np.ones(len(key), np.uint8), [key], world)[d]
                    w_pay.append(pay[ix])
                w_pay = np.concatenate(w_pay)
                _, _, ops, cols = got[d][name][step]
                assert len(ops) == len(w_pay) and np.array_equal(cols[1], w_pay), (name, step, d)


def _flat_setup(world, types, cap):
    import torch
    from risingwave_b200 import device
    total, ops_off, col_off = device.flat_layout(types, cap)
    bufs = [torch.zeros(total, dtype=torch.uint8, device="cuda") for _ in range(world)]
    flags = [torch.zeros(1024, dtype=torch.uint8, device="cuda") for _ in range(world)]
    views = []
    for b in bufs:
        ops = b[ops_off:ops_off + cap]
        cols = [b[o:o + cap * abi.TYPE_WIDTH[t]].view(device.TORCH_DTYPE[t]) for o, t in zip(col_off, types)]
        views.append((ops, cols))
    return bufs, flags, views


def test_flat_exchange_self_peer_feeds_counted_join(cuda, oracle):
    """world = 1 through the ONE-KERNEL exchange (rwgpu_shuffle_exchange_flat_device) into the counted join push, row count
    read on the device, the join reading the receive buffer IN PLACE -- the chain bench.py times at N>1 -- vs the oracle."""
    import torch
    from collections import Counter
    from risingwave_b200 import device, exchange
    from risingwave_b200.executor import HashJoinExecutor, JoinParams, MockSource
    from risingwave_b200.stream_chunk import net_multiset
    rng = np.random.default_rng(12)
    types = [abi.T_INT64] * 4
    nb, n, cap = 3000, 20000, 20000
    exs = []
    for be in (cuda, oracle):
        _, sl = MockSource.channel()
        _, sr = MockSource.channel()
        exs.append(HashJoinExecutor(be, abi.JOIN_INNER, sl.into_executor(types, [1]), sr.into_executor(types, [0]),
                                    JoinParams([0], [1]), JoinParams([0], []), [False], capacity_hint=nb))
    auct = [np.arange(nb, dtype=np.int64)] + [rng.integers(0, 1000, nb).astype(np.int64) for _ in range(3)]
    a_ops = np.full(nb, abi.OP_INSERT, np.uint8)
    for ex in exs:
        assert ex.eq_join_oneside(1, StreamChunk(a_ops, [Column(abi.T_INT64, c) for c in auct])) == []
    bufs, flags, views = _flat_setup(1, types, cap)
    counts = torch.zeros(1, dtype=torch.int64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    total_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
    total_host = torch.zeros(1, dtype=torch.int64).pin_memory()
    call = device.FlatExchangeCall([0], exchange.vnode_to_dest_table(1).cuda(), 1, 0, [bufs[0].data_ptr()], [flags[0].data_ptr()], cap,
                                   counts, err, total_dev.data_ptr(), total_host)
    stream = torch.cuda.Stream()
    for epoch in (1, 2, 3):
        m = n - 1000 * epoch
        bid = [rng.integers(0, nb + 200, m).astype(np.int64), (np.arange(m) + 10 ** 6 * epoch).astype(np.int64),
               rng.integers(0, 1 << 30, m).astype(np.int64), rng.integers(0, 1 << 30, m).astype(np.int64)]
        ops = np.full(m, abi.OP_INSERT, np.uint8)
        ops[rng.integers(0, m, 40)] = 0
        chunk = device.DeviceChunk(torch.from_numpy(ops).cuda(), [torch.from_numpy(c).cuda() for c in bid], types)
        with torch.cuda.stream(stream):
            call(chunk, epoch, stream)
            view = device.join_push_device(exs[0], abi.SIDE_LEFT, device.DeviceChunk(views[0][0], views[0][1], types), stream,
                                           n_rows_dev=total_dev.data_ptr())
            stream.synchronize()
        keep = ops != 0
        assert int(err.item()) == 0 and int(total_host.item()) == int(keep.sum()) == int(total_dev.item())
        # the receive buffer holds the visible rows in input order
        assert np.array_equal(views[0][1][1][:int(keep.sum())].cpu().numpy(), bid[1][keep])
        vis = view.visible()
        go, gc = view.ops().cpu().numpy(), [view.column(k).cpu().numpy() for k in range(view.n_cols)]
        if vis is not None:
            v = vis.cpu().numpy()
            go, gc = go[v], [c[v] for c in gc]
        got = Counter()
        for i in range(len(go)):
            got[tuple(int(c[i]) for c in gc)] += 1 if go[i] in (abi.OP_INSERT, abi.OP_UPDATE_INSERT) else -1
        want = net_multiset(exs[1].eq_join_oneside(0, StreamChunk(ops[keep], [Column(abi.T_INT64, c[keep]) for c in bid])))
        assert {k: v for k, v in got.items() if v} == dict(want), f"epoch {epoch}"


def test_flat_exchange_virtual_ranks(cuda, oracle):
    """flat_exchange_kernel with W virtual ranks on ONE device, one stream per rank, the W kernels running side
    by side (grids capped so they are co-resident) and meeting in the kernel's own cross-rank barriers: every receive buffer
    must hold the numpy stable partition, sources concatenated in rank order, with no gaps.  Two batches, so the second
    one reuses the flags with larger barrier values."""
    import torch
    from risingwave_b200 import device, exchange
    rng = np.random.default_rng(21)
    # (int64 x3: the staged path -- tiles partitioned in shared memory, coalesced stores; with an int32 column: row by row)
    for types, world, n in (([abi.T_INT64, abi.T_INT64, abi.T_INT32], 2, 70001), ([abi.T_INT64] * 3, 2, 70001), ([abi.T_INT64] * 3, 4, 9000),
                            ([abi.T_INT64, abi.T_INT64, abi.T_INT32], 4, 9000)):
        cap = world * n
        bufs, flags, views = _flat_setup(world, types, cap)
        v2d = exchange.vnode_to_dest_table(world).cuda()
        peers, flag_ptrs = [b.data_ptr() for b in bufs], [f.data_ptr() for f in flags]
        streams = [torch.cuda.Stream() for _ in range(world)]
        state = [dict(counts=torch.zeros(world, dtype=torch.int64, device="cuda"), err=torch.zeros(1, dtype=torch.int32, device="cuda"),
                      total=torch.zeros(1, dtype=torch.int64, device="cuda")) for _ in range(world)]
        calls = [device.FlatExchangeCall([0], v2d, world, r, peers, flag_ptrs, cap, state[r]["counts"], state[r]["err"],
                                         state[r]["total"].data_ptr(), None, max_blocks=24) for r in range(world)]
        for epoch in (1, 2):
            src, chunks = [], []
            for r in range(world):
                key = rng.integers(0, 100000, n).astype(np.int64)
                if epoch == 2 and r == 0:
                    key[:] = 5  # one source sends everything to one destination
                pay = np.arange(n, dtype=np.int64) + r * 10 ** 9 + epoch * 10 ** 7
                small = rng.integers(0, 100, n).astype(np.int32 if types[2] == abi.T_INT32 else np.int64)
                ops = rng.integers(1, 5, n).astype(np.uint8)
                ops[rng.random(n) < 0.05] = 0
                src.append((ops, [key, pay, small]))
                chunks.append(device.DeviceChunk(torch.from_numpy(ops).cuda(), [torch.from_numpy(c).cuda() for c in (key, pay, small)], types))
            torch.cuda.synchronize()
            for r in range(world):
                calls[r](chunks[r], epoch, streams[r])
            torch.cuda.synchronize()
            errs = [int(state[r]["err"].item()) for r in range(world)]
            if any(e & 2 for e in errs):
                pytest.skip("this device did not run the ranks' kernels side by side (barrier timed out)")
            assert errs == [0] * world
            for d in range(world):
                w_ops, w_cols = [], [[] for _ in types]
                for r in range(world):
                    ops, cols = src[r]
                    ix = _expected_partition(oracle, ops, [cols[0]], world)[d]
                    assert int(state[r]["counts"][d].item()) == len(ix)
                    w_ops.append(ops[ix])
                    for k in range(len(types)):
                        w_cols[k].append(cols[k][ix])
                w_ops = np.concatenate(w_ops)
                m = int(state[d]["total"].item())
                assert m == len(w_ops)
                assert np.array_equal(views[d][0][:m].cpu().numpy(), w_ops)
                for k in range(len(types)):
                    assert np.array_equal(views[d][1][k][:m].cpu().numpy(), np.concatenate(w_cols[k]))
