"""GPU: state persistence -- snapshot / restore round trips of both operators: an operator rebuilt
from the snapshot of another one emits exactly what the original emits on the rest of the stream, and the snapshot rows
themselves are what the oracle's state would be (the rows the reference keeps in its StateTables)."""
import numpy as np
import pytest

from risingwave_b200 import abi
from risingwave_b200.executor import AggCall, HashAggExecutor, HashJoinExecutor, JoinParams, MockSource
from risingwave_b200.stream_chunk import StreamChunk, concat_chunks, net_multiset

from helpers import JOIN_TYPES, make_agg
from test_gpu_join import StreamGen, StreamGen4

pytestmark = pytest.mark.gpu


def _mk_join(be, jt, types, null_safe=(False,)):
    _, sl = MockSource.channel()
    _, sr = MockSource.channel()
    return HashJoinExecutor(be, jt, sl.into_executor(types, [1]), sr.into_executor(types, [1]), JoinParams([0], [1]), JoinParams([0], [1]),
                            list(null_safe))


@pytest.mark.parametrize("name", ["Inner", "LeftOuter", "FullOuter", "LeftSemi", "RightAnti"])
@pytest.mark.parametrize("ncols", [4, 3])
def test_join_snapshot_restore_round_trip(cuda, oracle, name, ncols):
    jt = JOIN_TYPES[name]
    types = [abi.T_INT64] * ncols
    gen = (StreamGen4 if ncols == 4 else StreamGen)(seed=3 + jt, key_range=40, null_frac=0.05 if ncols == 3 else 0.0)
    a, o = _mk_join(cuda, jt, types), _mk_join(oracle, jt, types)
    first = [(int(gen.rng.integers(2)), gen.chunk(int(i % 2), int(gen.rng.integers(100, 900)), types=types)) for i in range(10)]
    first = [(i % 2, ch) for i, (_, ch) in enumerate(first)]
    for side, ch in first:
        assert net_multiset(a.eq_join_oneside(side, ch)) == net_multiset(o.eq_join_oneside(side, ch))
    a.flush_data(1)
    # the snapshot holds exactly the live rows of each side (what the reference's StateTable holds): compare with the generator's view
    snaps = [a.snapshot(s) for s in (0, 1)]
    for s in (0, 1):
        got = sorted(tuple(r) for c in snaps[s] for _, r in c.rows())
        want = sorted(r for r in gen.live[s] if not (r[0] is None))  # NULL-key rows never match and are never stored (not null-safe)
        assert got == want, f"side {s}"
    # a fresh operator restored from the snapshot behaves like the original from here on
    b = _mk_join(cuda, jt, types)
    for s in (0, 1):
        if snaps[s]:
            b.restore(s, concat_chunks(snaps[s]))
    for i in range(8):
        side = i % 2
        ch = gen.chunk(side, int(gen.rng.integers(100, 700)), p_delete=0.35, p_update=0.2, types=types)
        ga, gb, go = a.eq_join_oneside(side, ch), b.eq_join_oneside(side, ch), o.eq_join_oneside(side, ch)
        assert net_multiset(ga) == net_multiset(go), f"push {i}: original vs oracle"
        assert net_multiset(gb) == net_multiset(go), f"push {i}: restored vs oracle"


def test_agg_snapshot_restore_round_trip(cuda, oracle):
    cfgs = [
        {"schema": "III", "group_keys": [0], "agg_calls": ["(count:int8)", "(sum:int8 $1:int8)", "(max:int8 $2:int8)"], "append_only": True,
         "row_count_index": 0},
        {"schema": "IiF", "group_keys": [0, 1], "agg_calls": ["(count:int8)", "(sum:float8 $2:float8)", "(count:int8 $2:float8)"], "append_only": False,
         "row_count_index": 0},
        {"schema": "III", "group_keys": [0], "agg_calls": ["(count:int8)", "(min:int8 $1:int8)", "(max:int8 $2:int8)", "(sum:decimal $1:int8)"],
         "append_only": False, "row_count_index": 0},
    ]
    TY = {"I": abi.T_INT64, "i": abi.T_INT32, "F": abi.T_FLOAT64}
    for ci, cfg in enumerate(cfgs):
        types = [TY[ch] for ch in cfg["schema"]]
        (_, a), (_, o) = make_agg(cuda, cfg), make_agg(oracle, cfg)
        rng = np.random.default_rng(7 + ci)
        live = []

        def chunk(n, p_del):
            rows = []
            while len(rows) < n:
                if live and not cfg["append_only"] and rng.random() < p_del:
                    rows.append((abi.OP_DELETE, live.pop(int(rng.integers(len(live))))))
                else:
                    row = []
                    for k, t in enumerate(types):
                        v = int(rng.integers(0, 30 if k in cfg["group_keys"] else 50))
                        if k not in cfg["group_keys"] and rng.random() < 0.1:
                            v = None
                        elif t == abi.T_FLOAT64:
                            v = v / 8
                        row.append(v)
                    rows.append((abi.OP_INSERT, tuple(row)))
                    live.append(tuple(row))
            return StreamChunk.from_rows(types, rows)

        for epoch in range(3):
            ch = chunk(2000, 0.3)
            a.apply_chunk(ch)
            o.apply_chunk(ch)
            assert net_multiset(a.flush_data(epoch + 1)) == net_multiset(o.flush_data(epoch + 1))
        states, minput = a.snapshot()
        # the state rows ARE the operator's current output rows (a value state's output is its state datum)
        b_tx, b = make_agg(cuda, cfg)
        b.restore(concat_chunks(states), concat_chunks(minput) if minput else None)
        for epoch in range(3, 7):
            ch = chunk(1500, 0.5)
            a.apply_chunk(ch)
            b.apply_chunk(ch)
            o.apply_chunk(ch)
            ga, gb, go = a.flush_data(epoch + 1), b.flush_data(epoch + 1), o.flush_data(epoch + 1)
            assert net_multiset(ga) == net_multiset(go), f"cfg {ci} epoch {epoch}: original vs oracle"
            assert net_multiset(gb) == net_multiset(go), f"cfg {ci} epoch {epoch}: restored vs oracle"


@pytest.mark.parametrize("shape", ["unified", "two_keys"])
def test_watermark_state_cleaning(cuda, oracle, shape):
    """HashJoinExecutor::handle_watermark -> JoinHashMap::update_watermark (hash_join.rs:791-891): after a join-key
    watermark both sides drop every row below it at the next barrier.  The state shrinks (snapshot), and -- the
    watermark contract: no later row is below it -- every later result is still exactly the oracle's (which never cleans)."""
    if shape == "unified":
        types, keys, pk = [abi.T_INT64] * 3, [0], [1]
    else:
        types, keys, pk = [abi.T_INT64, abi.T_INT32, abi.T_INT64], [0, 1], [2]
    rng = np.random.default_rng(17)

    def mk(be):
        _, sl = MockSource.channel()
        _, sr = MockSource.channel()
        return HashJoinExecutor(be, abi.JOIN_INNER, sl.into_executor(types, pk), sr.into_executor(types, pk), JoinParams(keys, pk), JoinParams(keys, pk),
                                [False] * len(keys))

    g, o = mk(cuda), mk(oracle)
    next_pk = [0]

    def chunk(lo, hi, n):
        rows = []
        for _ in range(n):
            next_pk[0] += 1
            if shape == "unified":
                rows.append((abi.OP_INSERT, (int(rng.integers(lo, hi)), next_pk[0], int(rng.integers(0, 100)))))
            else:
                rows.append((abi.OP_INSERT, (int(rng.integers(lo, hi)), int(rng.integers(0, 3)), next_pk[0])))
        return StreamChunk.from_rows(types, rows)

    for i in range(6):  # event-time keys 0 .. 1000
        side = i % 2
        ch = chunk(0, 1000, 800)
        assert net_multiset(g.eq_join_oneside(side, ch)) == net_multiset(o.eq_join_oneside(side, ch))
    before = [sum(c.cardinality() for c in g.snapshot(s)) for s in (0, 1)]
    for s in (0, 1):
        g.update_watermark(s, 0, 600)
    g.flush_data(1)
    snaps = [g.snapshot(s) for s in (0, 1)]
    for s in (0, 1):
        rows = [r for c in snaps[s] for _, r in c.rows()]
        assert rows and all(r[0] >= 600 for r in rows)
        assert len(rows) < before[s] * 0.6
    for i in range(6):  # later rows respect the watermark
        side = i % 2
        ch = chunk(600, 1400, 700)
        assert net_multiset(g.eq_join_oneside(side, ch)) == net_multiset(o.eq_join_oneside(side, ch)), f"push {i} after cleaning"
