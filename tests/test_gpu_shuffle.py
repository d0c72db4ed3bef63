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
