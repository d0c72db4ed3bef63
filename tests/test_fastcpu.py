"""CPU: the timed CPU baseline (oracle/fastcpu.cc) computes the same results as the oracle."""
import ctypes as C
import os
import sys

import numpy as np

from risingwave_b200 import abi
from risingwave_b200.executor import AggCall, HashAggExecutor, HashJoinExecutor, JoinParams, MockSource
from risingwave_b200.stream_chunk import Column, StreamChunk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

M64 = (1 << 64) - 1


def test_join_matches_oracle(oracle):
    fc = bench.FastCpu().f
    fc.rwf_join_checksum.restype = C.c_uint64
    fc.rwf_join_checksum.argtypes = [C.c_void_p]
    h = fc.rwf_join_new()
    T4 = [abi.T_INT64] * 4
    _, sl = MockSource.channel()
    _, sr = MockSource.channel()
    ex = HashJoinExecutor(oracle, abi.JOIN_INNER, sl.into_executor(T4, [1]), sr.into_executor(T4, [0]),
                          JoinParams([0], [1]), JoinParams([0], []), [False])
    rng = np.random.default_rng(0)
    auct = bench.gen_auctions(3000, 1)
    bids = bench.gen_bids(5000, 0, 1, 3000)
    pushes = [(1, np.full(3000, 1, np.uint8), auct), (0, np.full(5000, 1, np.uint8), bids)]
    # retract some bids, update some auctions (U-/U+)
    sel = rng.choice(5000, 700, replace=False)
    pushes.append((0, np.full(700, 2, np.uint8), [c[sel] for c in bids]))
    ua = rng.choice(3000, 200, replace=False)
    ops = np.tile(np.array([4, 3], np.uint8), 200)
    cols = []
    for k, c in enumerate(auct):
        old = c[ua]
        new = old if k != 3 else old + 1
        cols.append(np.stack([old, new], 1).reshape(-1))
    pushes.append((1, ops, cols))
    want_rows, want_sum = 0, 0
    got_rows = 0
    for side, ops, cols in pushes:
        cols = [np.ascontiguousarray(c) for c in cols]
        got_rows += fc.rwf_join_push(h, side, len(ops), ops.ctypes.data, *[c.ctypes.data for c in cols])
        outs = ex.eq_join_oneside(side, StreamChunk(ops, [Column(abi.T_INT64, c) for c in cols]))
        for o in outs:
            for op, row in o.rows():
                want_rows += 1
                v = sum(w * c for w, c in zip(bench.CHECKSUM_WEIGHTS, row)) & M64
                want_sum = (want_sum + (v if op == abi.OP_INSERT else -v)) & M64
    assert got_rows == want_rows and want_rows > 5000
    assert fc.rwf_join_checksum(h) == want_sum
    fc.rwf_join_free(h)


def test_agg_matches_oracle(oracle):
    fc = bench.FastCpu().f
    fc.rwf_agg_checksum.restype = C.c_uint64
    fc.rwf_agg_checksum.argtypes = [C.c_void_p]
    a = fc.rwf_agg_new(1)
    _, src = MockSource.channel()
    ex = HashAggExecutor(oracle, src.into_executor([abi.T_INT64] * 2, []), True,
                         [AggCall.from_pretty(c) for c in ("(count:int8)", "(sum:int8 $1:int8)", "(max:int8 $1:int8)")], 0, [0])
    want_rows, want_sum, got_rows = 0, 0, 0
    for e in range(4):
        k, p = bench.gen_agg_rows(3000, e * 3000, 5)
        k = k % 500
        ops = np.full(3000, 1, np.uint8)
        fc.rwf_agg_push(a, 3000, ops.ctypes.data, k.ctypes.data, p.ctypes.data)
        got_rows += fc.rwf_agg_flush(a)
        ex.apply_chunk(StreamChunk(ops, [Column(abi.T_INT64, k), Column(abi.T_INT64, p)]))
        for o in ex.flush_data(e + 1):
            for op, row in o.rows():
                want_rows += 1
                v = sum(row) & M64
                want_sum = (want_sum + (v if op in (abi.OP_INSERT, abi.OP_UPDATE_INSERT) else -v)) & M64
    assert got_rows == want_rows
    assert fc.rwf_agg_checksum(a) == want_sum
    fc.rwf_agg_free(a)


def test_numpy_vnode_matches_oracle(oracle):
    keys = np.random.default_rng(1).integers(-2**62, 2**62, 2000).astype(np.int64)
    ch = StreamChunk(np.full(2000, 1, np.uint8), [Column(abi.T_INT64, keys)])
    assert np.array_equal(bench.vnode_of_int64(keys), oracle.vnode_compute(ch, [0], 256).astype(np.int32))
