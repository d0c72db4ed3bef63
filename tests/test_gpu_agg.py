"""GPU: HashAgg CUDA path (through the C ABI) vs the CPU oracle and the reference's golden vectors."""
import math

import numpy as np
import pytest

from risingwave_b200 import abi
from risingwave_b200.executor import AggCall, HashAggExecutor, MockSource
from risingwave_b200.stream_chunk import Column, StreamChunk, net_multiset

from helpers import load_golden, make_agg, rand_chunk, run_agg_kat

pytestmark = pytest.mark.gpu

AGG_KATS = [k for k in load_golden("hash_agg_kats.json") if "skipped" not in k]  # incl. test_hash_agg_min (retractable min)


@pytest.mark.parametrize("kat", AGG_KATS, ids=[k["name"] for k in AGG_KATS])
def test_hash_agg_golden(cuda, kat):
    run_agg_kat(cuda, kat)


def test_retractable_min_max_random_streams(cuda, oracle):
    """retractable min / max (AggState::MaterializedInput, agg_state.rs:49-56 / minput.rs) on the device: inserts,
    retractions of arbitrary stored values (incl. the current extreme and duplicates of it), NULL arguments, float
    arguments, groups that empty out and come back, several barriers -- vs the oracle's multiset restatement."""
    cfgs = [
        {"schema": "III", "group_keys": [0], "agg_calls": ["(count:int8)", "(min:int8 $1:int8)", "(max:int8 $2:int8)", "(sum:int8 $1:int8)"],
         "append_only": False, "row_count_index": 0},
        {"schema": "IFi", "group_keys": [0], "agg_calls": ["(count:int8)", "(max:float8 $1:float8)", "(min:int4 $2:int4)"],
         "append_only": False, "row_count_index": 0},
        {"schema": "IiI", "group_keys": [0, 1], "agg_calls": ["(count:int8)", "(max:int8 $2:int8)"], "append_only": False, "row_count_index": 0},
    ]
    TY = {"I": abi.T_INT64, "i": abi.T_INT32, "F": abi.T_FLOAT64}
    for ci, cfg in enumerate(cfgs):
        types = [TY[ch] for ch in cfg["schema"]]
        (_, ex_g), (_, ex_o) = make_agg(cuda, cfg), make_agg(oracle, cfg)
        rng = np.random.default_rng(90 + ci)
        live = []  # rows currently in the input
        store_g, store_o = Store(len(cfg["group_keys"])), Store(len(cfg["group_keys"]))
        for epoch in range(8):
            for _ in range(3):
                rows = []
                n = int(rng.integers(50, 400))
                while len(rows) < n:
                    if live and rng.random() < (0.6 if epoch % 3 == 2 else 0.3):
                        rows.append((abi.OP_DELETE, live.pop(int(rng.integers(len(live))))))
                    else:
                        row = []
                        for k, t in enumerate(types):
                            hi = 6 if k in cfg["group_keys"] else 12  # few distinct values: duplicates of the extreme
                            v = None if (k not in cfg["group_keys"] and rng.random() < 0.1) else int(rng.integers(0, hi))
                            if t == abi.T_FLOAT64 and v is not None:
                                v = v / 4 - 1.0
                            row.append(v)
                        row = tuple(row)
                        rows.append((abi.OP_INSERT, row))
                        live.append(row)
                ch = StreamChunk.from_rows(types, rows)
                ex_g.apply_chunk(ch)
                ex_o.apply_chunk(ch)
            g, o = ex_g.flush_data(epoch + 1), ex_o.flush_data(epoch + 1)
            store_g.apply(g)
            store_o.apply(o)
            assert store_g.rows == store_o.rows, f"cfg {ci} epoch {epoch}"
            assert net_multiset(g) == net_multiset(o), f"cfg {ci} epoch {epoch}"


def test_retracting_a_value_that_was_never_inserted_is_inconsistent(cuda):
    cfg = {"schema": "II", "group_keys": [0], "agg_calls": ["(count:int8)", "(min:int8 $1:int8)"], "append_only": False, "row_count_index": 0}
    _, ex = make_agg(cuda, cfg)
    ex.apply_chunk(StreamChunk.from_pretty(" I I\n + 1 5\n + 1 7"))
    ex.flush_data(1)
    ex.apply_chunk(StreamChunk.from_pretty(" I I\n - 1 6"))
    with pytest.raises(abi.RwError) as e:
        ex.flush_data(2)
    assert e.value.code == abi.RW_ERR_INCONSISTENT


class Store:
    """multiset-free keyed store: apply deltas, checking every delete against what was inserted
    (the `Store::apply_chunk` idea of snapshot.rs:219-254, keyed by group key)."""

    def __init__(self, n_keys):
        self.n_keys = n_keys
        self.rows = {}

    def apply(self, chunks):
        for ch in chunks:
            for op, row in ch.rows():
                k = row[: self.n_keys]
                if op in (abi.OP_INSERT, abi.OP_UPDATE_INSERT):
                    assert k not in self.rows, f"double insert of group {k}"
                    self.rows[k] = row
                else:
                    assert k in self.rows, f"delete of missing group {k}"
                    old = self.rows.pop(k)
                    assert _rows_close(old, row, 0.0), f"delete row {row} != stored {old}"


def _rows_close(a, b, rtol):
    for x, y in zip(a, b):
        if x is None or y is None:
            if x is not y:
                return False
        elif isinstance(x, float) or isinstance(y, float):
            if math.isnan(x) and math.isnan(y):
                continue
            if x != y and not (abs(x - y) <= rtol * max(abs(x), abs(y))):
                return False
        elif x != y:
            return False
    return True


def run_both(cuda, oracle, cfg, epochs, exact=True, rtol=0.0, chunk_size=1024, hint=0):
    """epochs: list of lists of StreamChunk.  Compares per-epoch applied state."""
    stores = []
    outs = []
    for be in (cuda, oracle):
        tx, src = MockSource.channel()
        src = src.into_executor(cfg["types"], [])
        ex = HashAggExecutor(be, src, cfg["append_only"], [AggCall.from_pretty(c) for c in cfg["calls"]], 0,
                             cfg["keys"], chunk_size, group_capacity_hint=hint if be is cuda else 0)
        st = Store(len(cfg["keys"]))
        per_epoch = []
        for e, chunks in enumerate(epochs):
            for ch in chunks:
                ex.apply_chunk(ch)
            out = ex.flush_data(e + 1)
            for oc in out:  # chunk cut rule: <= chunk_size (+1 when a U- would be last)
                assert oc.capacity() <= chunk_size + 1
                assert oc.ops[-1] != abi.OP_UPDATE_DELETE
            st.apply(out)
            per_epoch.append((net_multiset(out), dict(st.rows)))
        outs.append(per_epoch)
        stores.append(st)
    for e, (g, o) in enumerate(zip(outs[0], outs[1])):
        if exact:
            assert g[0] == o[0], f"epoch {e}: emitted net delta differs"
        assert g[1].keys() == o[1].keys(), f"epoch {e}: group sets differ"
        for k in g[1]:
            assert _rows_close(g[1][k], o[1][k], rtol), f"epoch {e} group {k}: {g[1][k]} vs {o[1][k]}"
    return outs


def test_count_sum_max_append_only_random(cuda, oracle):
    rng = np.random.default_rng(1)
    cfg = dict(types=[abi.T_INT64] * 3, keys=[0], append_only=True,
               calls=["(count:int8)", "(sum:int8 $1:int8)", "(max:int8 $2:int8)", "(min:int8 $2:int8)"])
    epochs = [[rand_chunk(rng, 1024, cfg["types"], key_cols=(0,), key_range=200) for _ in range(3)] for _ in range(4)]
    run_both(cuda, oracle, cfg, epochs)


def test_retractions_nulls_invisible(cuda, oracle):
    """count/sum with Delete / UpdateDelete rows, NULL keys and args, invisible rows."""
    rng = np.random.default_rng(2)
    cfg = dict(types=[abi.T_INT64, abi.T_INT64, abi.T_INT32], keys=[0], append_only=False,
               calls=["(count:int8)", "(sum:int8 $1:int8)", "(count:int8 $2:int4)", "(sum:int8 $2:int4)", "(sum0:int8 $1:int8)"])
    epochs2 = []
    live = []
    rng = np.random.default_rng(3)
    for e in range(5):
        chunks = []
        for _ in range(2):
            n = 700
            base = rand_chunk(rng, n, cfg["types"], key_cols=(0,), key_range=50, null_frac=0.15)
            rows, ops, vis = [], [], []
            for i in range(n):
                if live and rng.random() < 0.35:
                    rows.append(live.pop(rng.integers(len(live))))
                    ops.append(abi.OP_DELETE if rng.random() < 0.5 else abi.OP_UPDATE_DELETE)
                    vis.append(True)
                else:
                    v = rng.random() < 0.9
                    rows.append(base.row(i))
                    ops.append(abi.OP_INSERT if rng.random() < 0.7 else abi.OP_UPDATE_INSERT)
                    vis.append(v)
                    if v:
                        live.append(base.row(i))
            ch = StreamChunk.from_rows(cfg["types"], list(zip(ops, rows)))
            ch.vis = np.array(vis, dtype=bool)
            chunks.append(ch)
        epochs2.append(chunks)
    # final epoch: delete everything that is still live -> every group emits a Delete
    rows = live
    ch = StreamChunk.from_rows(cfg["types"], [(abi.OP_DELETE, r) for r in rows])
    epochs2.append([ch])
    outs = run_both(cuda, oracle, cfg, epochs2)
    assert not outs[0][-1][1], "all groups must be gone after deleting every row"


def test_multi_column_and_narrow_keys(cuda, oracle):
    rng = np.random.default_rng(4)
    cfg = dict(types=[abi.T_INT64, abi.T_INT32, abi.T_INT16, abi.T_INT64], keys=[0, 1, 2], append_only=True,
               calls=["(count:int8)", "(sum:int8 $3:int8)", "(max:int8 $3:int8)"])
    epochs = [[rand_chunk(rng, 1500, cfg["types"], key_cols=(0, 1, 2), key_range=6, null_frac=0.1)] for _ in range(3)]
    run_both(cuda, oracle, cfg, epochs)


def test_float_sum_and_minmax(cuda, oracle):
    """sum(float8)/sum(float4) within 1e-6 relative (atomics reorder the additions; the reference
    adds in row order, general.rs:28-41); min/max on floats exact."""
    rng = np.random.default_rng(5)
    cfg = dict(types=[abi.T_INT32, abi.T_FLOAT64, abi.T_FLOAT32], keys=[0], append_only=True,
               calls=["(count:int8)", "(sum:float8 $1:float8)", "(sum:float4 $2:float4)", "(min:float8 $1:float8)",
                      "(max:float4 $2:float4)"])
    epochs = [[rand_chunk(rng, 2000, cfg["types"], key_cols=(0,), key_range=40, null_frac=0.05)] for _ in range(3)]
    run_both(cuda, oracle, cfg, epochs, exact=False, rtol=1e-6)


def test_sum_int8_to_decimal_and_overflow(cuda, oracle):
    big = 2**62
    cfg = dict(types=[abi.T_INT64, abi.T_INT64], keys=[0], append_only=False,
               calls=["(count:int8)", "(sum:decimal $1:int8)"])
    ch = StreamChunk.from_pretty(f" I I\n + 1 {big}\n + 1 {big}\n + 1 {big}\n + 2 -{big}\n + 2 -{big}\n + 2 -{big}\n + 3 7")
    outs = run_both(cuda, oracle, cfg, [[ch], [StreamChunk.from_pretty(f" I I\n - 1 {big}\n + 3 1")]])
    assert outs[0][0][1][(1,)] == (1, 3, 3 * big)
    # the internal sum(int8)->int8 form overflows => ExprError::NumericOutOfRange (general.rs:32-40)
    cfg2 = dict(cfg, calls=["(count:int8)", "(sum:int8 $1:int8)"])
    for be in (cuda, oracle):
        tx, src = MockSource.channel()
        ex = HashAggExecutor(be, src.into_executor(cfg2["types"], []), False, [AggCall.from_pretty(c) for c in cfg2["calls"]], 0, [0])
        with pytest.raises(abi.RwError) as e:  # the reference errors while applying, the GPU at the barrier
            ex.apply_chunk(ch)
            ex.flush_data(1)
        assert e.value.code == abi.RW_ERR_NUMERIC_OUT_OF_RANGE


def test_sentinel_and_extreme_keys(cuda, oracle):
    i64min, i64max = -2**63, 2**63 - 1
    cfg = dict(types=[abi.T_INT64, abi.T_INT64], keys=[0], append_only=True, calls=["(count:int8)", "(sum:int8 $1:int8)"])
    ch = StreamChunk.from_pretty(f" I I\n + {i64min} 1\n + {i64max} 2\n + . 3\n + 0 4\n + -1 5\n + {i64min} 6\n + . 7")
    run_both(cuda, oracle, cfg, [[ch], [ch]])


def test_table_growth_many_groups(cuda, oracle):
    """group_capacity_hint tiny => several rehashes while groups keep their state and prev outputs."""
    rng = np.random.default_rng(6)
    cfg = dict(types=[abi.T_INT64, abi.T_INT64], keys=[0], append_only=True,
               calls=["(count:int8)", "(sum:int8 $1:int8)", "(max:int8 $1:int8)"])
    epochs = [[rand_chunk(rng, 4096, cfg["types"], key_cols=(0,), key_range=20000) for _ in range(2)] for _ in range(4)]
    run_both(cuda, oracle, cfg, epochs, hint=16)


def test_negative_row_count_strict(cuda):
    tx, src = MockSource.channel()
    ex = HashAggExecutor(cuda, src.into_executor([abi.T_INT64, abi.T_INT64], []), False,
                         [AggCall.from_pretty("(count:int8)")], 0, [0])
    ex.apply_chunk(StreamChunk.from_pretty(" I I\n - 1 1"))
    with pytest.raises(abi.RwError) as e:
        ex.flush_data(1)
    assert e.value.code == abi.RW_ERR_INCONSISTENT


def test_empty_and_ragged_inputs(cuda, oracle):
    cfg = dict(types=[abi.T_INT64, abi.T_INT64], keys=[0], append_only=True, calls=["(count:int8)", "(sum:int8 $1:int8)"])
    empty = StreamChunk.from_pretty(" I I")
    one = StreamChunk.from_pretty(" I I\n + 1 1")
    allinv = StreamChunk.from_pretty(" I I\n + 1 1 D\n + 2 2 D")
    rng = np.random.default_rng(8)
    ragged = [rand_chunk(rng, n, cfg["types"], key_cols=(0,), key_range=9) for n in (1, 63, 64, 65, 127, 1025)]
    run_both(cuda, oracle, cfg, [[empty], [one, allinv], ragged, []])


def test_large_against_numpy(cuda):
    """BASELINE cfg2 sizes (2^20 rows / epoch, 2^20-key domain): size-independent properties --
    sum of per-group counts == rows, sum of sums == column sum, max of max == column max -- and an
    exact comparison with a numpy group-by."""
    rng = np.random.default_rng(9)
    n = 1 << 20
    keys = rng.integers(0, 1 << 20, n).astype(np.int64)
    price = rng.integers(0, 1 << 24, n).astype(np.int64)
    tx, src = MockSource.channel()
    ex = HashAggExecutor(cuda, src.into_executor([abi.T_INT64, abi.T_INT64], []), True,
                         [AggCall.from_pretty(c) for c in ("(count:int8)", "(sum:int8 $1:int8)", "(max:int8 $1:int8)")], 0, [0])
    for i in range(0, n, 1 << 16):
        sl = slice(i, i + (1 << 16))
        ex.apply_chunk(StreamChunk(np.full(1 << 16, abi.OP_INSERT, np.uint8), [Column(abi.T_INT64, keys[sl]), Column(abi.T_INT64, price[sl])]))
    out = ex.flush_data(1)
    k = np.concatenate([c.columns[0].data for c in out])
    cnt = np.concatenate([c.columns[1].data for c in out])
    sm = np.concatenate([c.columns[2].data for c in out])
    mx = np.concatenate([c.columns[3].data for c in out])
    assert all((c.ops == abi.OP_INSERT).all() for c in out)
    assert cnt.sum() == n and sm.sum() == price.sum() and mx.max() == price.max()
    order = np.argsort(k)
    uk, inv, ucnt = np.unique(keys, return_inverse=True, return_counts=True)
    assert np.array_equal(k[order], uk) and np.array_equal(cnt[order], ucnt)
    usum = np.bincount(inv, weights=None, minlength=len(uk)) * 0
    usum = np.zeros(len(uk), np.int64)
    np.add.at(usum, inv, price)
    umax = np.zeros(len(uk), np.int64)
    np.maximum.at(umax, inv, price)
    assert np.array_equal(sm[order], usum) and np.array_equal(mx[order], umax)
    # idempotence: a barrier with no input emits nothing
    assert ex.flush_data(2) == []


# ------------------------------------------------------------------------------------------ round 2: device-resident API
def _device_epochs(cuda, n_epochs, rows, keys, seed, hint):
    """q4-shaped input (count(*), sum, max GROUP BY key) as device chunks + the same rows for the oracle"""
    import torch
    from risingwave_b200 import device
    rng = np.random.default_rng(seed)
    eps = []
    for e in range(n_epochs):
        k = rng.integers(0, keys, rows).astype(np.int64)
        v = rng.integers(0, 1 << 24, rows).astype(np.int64)
        host = StreamChunk(np.full(rows, abi.OP_INSERT, np.uint8), [Column(abi.T_INT64, k), Column(abi.T_INT64, v)])
        dev = device.DeviceChunk(torch.ones(rows, dtype=torch.uint8, device="cuda"), [torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda()],
                                 [abi.T_INT64] * 2)
        eps.append((host, dev))
    return eps


def _view_net(view):
    """net applied change of a device delta: {row: #inserts - #deletes}, zero entries dropped"""
    from collections import Counter
    ops = view.ops().cpu().numpy()
    cols = [view.column(k).cpu().numpy() for k in range(view.n_cols)]
    c = Counter()
    for i in range(view.n_rows):
        row = tuple(int(col[i]) for col in cols)
        c[row] += 1 if ops[i] in (abi.OP_INSERT, abi.OP_UPDATE_INSERT) else -1
    return {k: v for k, v in c.items() if v}


def test_async_barriers_two_outstanding_match_oracle(cuda, oracle):
    """rwgpu_agg_flush_device_async / _collect: the delta of barrier e is collected while epoch e + 1 has already been
    pushed (and its barrier enqueued): two output sets.  Every collected delta equals the oracle's for that epoch."""
    import torch
    from risingwave_b200 import device
    calls = ["(count:int8)", "(sum:int8 $1:int8)", "(max:int8 $1:int8)"]
    exs = []
    for be in (cuda, oracle):
        _, src = MockSource.channel()
        exs.append(HashAggExecutor(be, src.into_executor([abi.T_INT64] * 2, []), True, [AggCall.from_pretty(c) for c in calls], 0, [0],
                                   group_capacity_hint=64))  # tiny hint: the table grows while barriers are outstanding
    eps = _device_epochs(cuda, 7, 5000, 3000, seed=4, hint=64)
    stream = torch.cuda.Stream()
    want = []
    for e, (host, _) in enumerate(eps):
        exs[1].apply_chunk(host)
        want.append(net_multiset(exs[1].flush_data(e + 1)))
    got = []
    with torch.cuda.stream(stream):
        for e, (_, dev) in enumerate(eps):
            device.agg_push_device(exs[0], dev, stream)
            device.agg_flush_device_async(exs[0], e + 1, stream)
            if e > 0:
                got.append(_view_net(device.agg_flush_collect(exs[0], stream)))
        got.append(_view_net(device.agg_flush_collect(exs[0], stream)))
        with pytest.raises(abi.RwError):
            device.agg_flush_collect(exs[0], stream)  # nothing outstanding
    assert len(got) == len(want)
    for e, (g, w) in enumerate(zip(got, want)):
        assert g == dict(w), f"epoch {e}"


def test_pushes_on_a_caller_stream_with_growth_inside_an_epoch(cuda, oracle):
    """several rwgpu_agg_push_device calls on a CALLER's stream inside one epoch, the later ones
    forcing the table to grow -- growth used to read the group count and re-hash on the handle's own stream,
    racing with the apply kernels still running on the caller's stream."""
    import torch
    from risingwave_b200 import device
    calls = ["(count:int8)", "(sum:int8 $1:int8)"]
    exs = []
    for be in (cuda, oracle):
        _, src = MockSource.channel()
        exs.append(HashAggExecutor(be, src.into_executor([abi.T_INT64] * 2, []), True, [AggCall.from_pretty(c) for c in calls], 0, [0],
                                   group_capacity_hint=16))
    stream = torch.cuda.Stream()
    for epoch in range(3):
        eps = _device_epochs(cuda, 6, 40000, 200000, seed=50 + epoch, hint=16)
        with torch.cuda.stream(stream):
            for host, dev in eps:
                device.agg_push_device(exs[0], dev, stream)  # no sync in between
                exs[1].apply_chunk(host)
            g = _view_net(device.agg_flush_device(exs[0], epoch + 1, stream))
        assert g == dict(net_multiset(exs[1].flush_data(epoch + 1))), f"epoch {epoch}"
