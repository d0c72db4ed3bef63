"""Shared test helpers: run the reference's golden test scripts through an executor Backend."""
import json
import os

import numpy as np

from risingwave_b200 import abi
from risingwave_b200.executor import (AggCall, HashAggExecutor, HashJoinExecutor, JoinParams, MockSource, PENDING)
from risingwave_b200.stream_chunk import StreamChunk, PRETTY_TYPES, net_multiset, emitted_multiset

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

JOIN_TYPES = {"Inner": abi.JOIN_INNER, "LeftOuter": abi.JOIN_LEFT_OUTER, "RightOuter": abi.JOIN_RIGHT_OUTER,
              "FullOuter": abi.JOIN_FULL_OUTER, "LeftSemi": abi.JOIN_LEFT_SEMI, "LeftAnti": abi.JOIN_LEFT_ANTI,
              "RightSemi": abi.JOIN_RIGHT_SEMI, "RightAnti": abi.JOIN_RIGHT_ANTI}


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def schema_types(s):
    return [PRETTY_TYPES[ch] for ch in s]


def make_join(backend, cfg, chunk_size=1024):
    types = schema_types(cfg["schema"])
    tx_l, src_l = MockSource.channel()
    tx_r, src_r = MockSource.channel()
    src_l = src_l.into_executor(types, cfg["stream_key"])
    src_r = src_r.into_executor(types, cfg["stream_key"])
    ex = HashJoinExecutor(backend, JOIN_TYPES[cfg["join_type"]], src_l, src_r,
                          JoinParams(cfg["join_keys"], cfg["deduped_pk"]),
                          JoinParams(cfg["join_keys"], cfg["deduped_pk"]),
                          cfg["null_safe"], None, cfg.get("cond"), cfg["append_only"], chunk_size)
    return tx_l, tx_r, ex


def run_join_kat(backend, kat, exact=True):
    """Replays one hash_join.rs test.  exact=True: StreamChunk equality incl. visibility (what the
    reference asserts).  exact=False: per-step net applied multiset (order-insensitive parity)."""
    tx_l, tx_r, ex = make_join(backend, kat["config"])
    stream = ex.execute()
    tx = {"l": tx_l, "r": tx_r}
    for i, st in enumerate(kat["steps"]):
        op = st["op"]
        if op == "push_chunk":
            tx[st["side"]].push_chunk(StreamChunk.from_pretty(st["chunk"]))
        elif op == "push_barrier":
            tx[st["side"]].push_barrier(st["epoch"])
        elif op == "expect_pending":
            if exact:
                stream.next_unwrap_pending()
            else:  # a chunk whose rows are all invisible / net-zero is also acceptable
                m = stream.poll_next()
                if m is not PENDING:
                    assert m.chunk is not None and not net_multiset([m.chunk]), f"step {i}: expected pending, got {m}"
                    stream.next_unwrap_pending()
        elif op == "expect_barrier":
            stream.next_unwrap_ready_barrier()
        elif op == "expect_chunk":
            want = StreamChunk.from_pretty(st["chunk"])
            got = stream.next_unwrap_ready_chunk()
            if exact:
                if st.get("compact_vis"):
                    got = StreamChunk.from_rows(got.types(), list(got.rows()))
                assert got == want, f"{kat['name']} step {i}:\n got\n{got}\n want\n{want}"
            else:
                assert net_multiset([got]) == net_multiset([want]), \
                    f"{kat['name']} step {i}:\n got\n{got}\n want\n{want}"
        else:
            raise AssertionError(op)


def make_agg(backend, cfg, chunk_size=1024):
    tx, src = MockSource.channel()
    src = src.into_executor(schema_types(cfg["schema"]), [])
    calls = [AggCall.from_pretty(c) for c in cfg["agg_calls"]]
    ex = HashAggExecutor(backend, src, cfg["append_only"], calls, cfg["row_count_index"], cfg["group_keys"], chunk_size)
    return tx, ex


REF_OP_ORDER = {abi.OP_INSERT: 0, abi.OP_DELETE: 1, abi.OP_UPDATE_DELETE: 2, abi.OP_UPDATE_INSERT: 3}
OPTOK = {"+": abi.OP_INSERT, "-": abi.OP_DELETE, "U-": abi.OP_UPDATE_DELETE, "U+": abi.OP_UPDATE_INSERT}


def sorted_rows(chunk):
    """Op derive(Ord) order Insert < Delete < UpdateDelete < UpdateInsert, then row (snapshot.rs sort_chunk)."""
    return sorted(((REF_OP_ORDER[op], tuple((v is None, 0 if v is None else v) for v in row)) for op, row in chunk.rows()))


def run_agg_kat(backend, kat):
    tx, ex = make_agg(backend, kat["config"])
    stream = ex.execute()
    for st in kat["steps"]:
        if st["op"] == "push_chunk":
            tx.push_chunk(StreamChunk.from_pretty(st["chunk"]))
        else:
            tx.push_barrier(st["epoch"])
    msgs = stream.drain_until_pending()
    exp = kat["expected"]
    assert len(msgs) == len(exp), (msgs, exp)
    for m, e in zip(msgs, exp):
        if "barrier" in e:
            assert m.barrier is not None and m.barrier.epoch == e["barrier"]
        else:
            want = sorted((REF_OP_ORDER[OPTOK[r[0]]], tuple((False, int(x)) for x in r[1:])) for r in e["chunk_rows"])
            assert sorted_rows(m.chunk) == want, f"{kat['name']}: got\n{m.chunk}\nwant {e['chunk_rows']}"


def rand_chunk(rng, n, types, key_cols=(), key_range=16, null_frac=0.0, ops=None, vis_frac=1.0):
    """random StreamChunk of n rows; key columns drawn from a small range to force collisions."""
    from risingwave_b200.stream_chunk import Column, NP_DTYPE
    cols = []
    for k, t in enumerate(types):
        if t in (abi.T_FLOAT32, abi.T_FLOAT64):
            data = rng.integers(-1000, 1000, n).astype(NP_DTYPE[t]) / 4
        else:
            hi = key_range if k in key_cols else 1000
            data = rng.integers(0, hi, n).astype(NP_DTYPE[t])
        valid = None
        if null_frac > 0:
            valid = rng.random(n) >= null_frac
        cols.append(Column(t, data, valid))
    if ops is None:
        ops = np.full(n, abi.OP_INSERT, np.uint8)
    vis = None if vis_frac >= 1.0 else rng.random(n) < vis_frac
    return StreamChunk(ops, cols, vis)


def run_nexmark_q4(oracle):
    """The reference's own SQL-level fixture for q4 (e2e_test/nexmark/insert_{auction,bid}.slt.part ->
    e2e_test/streaming/nexmark/views/q4.slt.part -> expected rows of e2e_test/streaming/nexmark/q4.slt.part),
    run incrementally through the oracle's operators exactly as the streaming plan does:
    bid JOIN auction ON auction = id -> Filter(date_time BETWEEN a.date_time AND a.expires) -> Project ->
    HashAgg(max(price) GROUP BY id, category) -> Project -> HashAgg(count, sum GROUP BY category) -> avg.
    Several barriers, so the first aggregation retracts and re-emits groups whose max changed."""
    from decimal import Decimal
    from fractions import Fraction
    from risingwave_b200.executor import AggCall, FilterExecutor, HashAggExecutor, HashJoinExecutor, JoinParams, MockSource
    from risingwave_b200.stream_chunk import Column, StreamChunk
    fx = load_golden("nexmark_q4_fixture.json")
    I = abi.T_INT64
    _, sl = MockSource.channel()
    _, sr = MockSource.channel()
    join = HashJoinExecutor(oracle, abi.JOIN_INNER, sl.into_executor([I] * 4, [3]), sr.into_executor([I] * 4, [0]),
                            JoinParams([0], [3]), JoinParams([0], [0]), [False])
    _, sf = MockSource.channel()
    flt = FilterExecutor(oracle, sf.into_executor([I] * 8, []),
                         "(and:boolean (greater_than_or_equal:boolean $2:int8 $5:int8) (less_than_or_equal:boolean $2:int8 $6:int8))")
    _, s1 = MockSource.channel()
    # the real plan's inner aggregation is NOT append-only (its input is a join's change stream): max(price) is a
    # retractable max = MaterializedInput state (nexmark.yaml q4 stream plan)
    agg1 = HashAggExecutor(oracle, s1.into_executor([I] * 3, []), False,
                           [AggCall.from_pretty(c) for c in ("(count:int8)", "(max:int8 $2:int8)")], 0, [0, 1])
    _, s2 = MockSource.channel()
    agg2 = HashAggExecutor(oracle, s2.into_executor([I] * 2, []), False,
                           [AggCall.from_pretty(c) for c in ("(count:int8)", "(sum:int8 $1:int8)")], 0, [0])
    mv = {}  # category -> (count, sum)

    def ins(rows):
        cols = list(zip(*rows))
        return StreamChunk(np.full(len(rows), abi.OP_INSERT, np.uint8), [Column(I, np.array(c, dtype=np.int64)) for c in cols])

    def after_join(chunks):
        for ch in chunks:
            f = flt.filter(ch)
            if f is not None:  # Project (a.id, a.category, b.price)
                agg1.apply_chunk(StreamChunk(f.ops, [f.columns[4], f.columns[7], f.columns[1]], f.vis))

    def barrier(epoch):
        for ch in agg1.flush_data(epoch):  # (id, category, count, max) -> Project (category, max)
            agg2.apply_chunk(StreamChunk(ch.ops, [ch.columns[1], ch.columns[3]], ch.vis))
        for ch in agg2.flush_data(epoch):
            for op, row in ch.rows():
                if op in (abi.OP_INSERT, abi.OP_UPDATE_INSERT):
                    mv[row[0]] = (row[1], row[2])
                elif mv.get(row[0]) == (row[1], row[2]):
                    del mv[row[0]]

    bids = [r + [k] for k, r in enumerate(fx["bid"])]  # + a row id as the stream key
    auct, epoch = fx["auction"], 0
    for step in range(5):  # interleave the two inputs, a barrier after every slice
        after_join(join.eq_join_oneside(1, ins(auct[step * 8:(step + 1) * 8])))
        after_join(join.eq_join_oneside(0, ins(bids[step * 10:(step + 1) * 10])))
        epoch += 1
        barrier(epoch)
    want = {int(c): Fraction(Decimal(v)) for c, v in fx["expected_q4"]}
    got = {c: Fraction(s, n) for c, (n, s) in mv.items()}
    assert got == want


def run_nexmark_q7(backend):
    """The reference's SQL-level fixture for Nexmark q7 (e2e_test/streaming/nexmark/views/q7.slt.part -> expected rows
    of e2e_test/streaming/nexmark/q7.slt.part), incrementally:
      bid -> [Project: + window_end = tumble_end(date_time, 10 s)] -> HashAgg(max(price) GROUP BY window_end)
          -> [Project: (maxprice, window_end, window_end - 10 s)] = B1
      bid B JOIN B1 ON B.price = B1.maxprice -> Filter(B.date_time BETWEEN B1.window_end - 10 s AND B1.window_end)
    The aggregate is re-emitted at every barrier, so the join's right side sees U-/U+ retractions of old maxima.
    (The two Projects carry arithmetic and stay on the host, as in the shim.)"""
    from collections import Counter
    from risingwave_b200.executor import AggCall, FilterExecutor, HashAggExecutor, HashJoinExecutor, JoinParams, MockSource
    from risingwave_b200.stream_chunk import Column, StreamChunk
    fx = load_golden("nexmark_q7_fixture.json")
    I = abi.T_INT64
    W = 10_000_000  # 10 s in microseconds
    _, sa = MockSource.channel()
    agg = HashAggExecutor(backend, sa.into_executor([I, I], []), True,
                          [AggCall.from_pretty(c) for c in ("(count:int8)", "(max:int8 $1:int8)")], 0, [0])
    _, sl = MockSource.channel()
    _, sr = MockSource.channel()
    # left: bid (auction, bidder, price, date_time, row id)   right: B1 (maxprice, window_end, window_start)
    join = HashJoinExecutor(backend, abi.JOIN_INNER, sl.into_executor([I] * 5, [4]), sr.into_executor([I] * 3, [1]),
                            JoinParams([2], [4]), JoinParams([0], [1]), [False])
    _, sf = MockSource.channel()
    flt = FilterExecutor(backend, sf.into_executor([I] * 8, []),
                         "(and:boolean (greater_than_or_equal:boolean $3:int8 $7:int8) (less_than_or_equal:boolean $3:int8 $6:int8))")
    mv = Counter()

    def sink(chunks):
        for ch in chunks:
            f = flt.filter(ch)
            if f is None:
                continue
            for op, row in f.rows():  # (visible rows only)
                key = (row[0], row[2], row[1], row[3])  # auction, price, bidder, date_time
                mv[key] += 1 if op in (abi.OP_INSERT, abi.OP_UPDATE_INSERT) else -1

    def ins(rows, ops=None):
        cols = list(zip(*rows))
        return StreamChunk(np.full(len(rows), abi.OP_INSERT, np.uint8) if ops is None else ops,
                           [Column(I, np.array(c, dtype=np.int64)) for c in cols])

    bids = [r + [k] for k, r in enumerate(fx["bid"])]
    for e, lo in enumerate(range(0, len(bids), 7)):
        part = bids[lo:lo + 7]
        sink(join.eq_join_oneside(0, ins(part)))
        agg.apply_chunk(ins([[(r[3] // W) * W + W, r[2]] for r in part]))  # (window_end, price)
        for ch in agg.flush_data(e + 1):  # (window_end, count, max) -> B1 rows, ops preserved (U-/U+ on a changed max)
            rows = [[row[2], row[0], row[0] - W] for _, row in ch.rows()]
            ops = np.array([op for op, _ in ch.rows()], dtype=np.uint8)
            if rows:
                sink(join.eq_join_oneside(1, ins(rows, ops)))
    got = sorted(k for k, v in mv.items() for _ in range(v))
    assert all(v >= 0 for v in mv.values())
    assert got == sorted(tuple(r) for r in fx["expected_q7"])


def run_nexmark_q8(backend):
    """The reference's SQL-level fixture for Nexmark q8 (e2e_test/streaming/nexmark/views/q8.slt.part -> expected rows
    of e2e_test/streaming/nexmark/q8.slt.part), incrementally:
      person  -> [Project: id, tumble window (start, end)] -> HashAgg(GROUP BY id, start, end) = P
      auction -> [Project: seller, tumble window]         -> HashAgg(GROUP BY seller, start, end) = A
      P JOIN A ON id = seller AND starttime = starttime AND endtime = endtime      (three-column join key)
    `name` is functionally dependent on `id` and is attached at the end (varchar columns stay on the host)."""
    from collections import Counter
    from risingwave_b200.executor import AggCall, HashAggExecutor, HashJoinExecutor, JoinParams, MockSource
    from risingwave_b200.stream_chunk import Column, StreamChunk
    fx = load_golden("nexmark_q8_fixture.json")
    I = abi.T_INT64
    W = 10_000_000
    aggs = []
    for _ in range(2):
        _, s = MockSource.channel()
        aggs.append(HashAggExecutor(backend, s.into_executor([I] * 3, []), True, [AggCall.from_pretty("(count:int8)")], 0, [0, 1, 2]))
    _, sl = MockSource.channel()
    _, sr = MockSource.channel()
    join = HashJoinExecutor(backend, abi.JOIN_INNER, sl.into_executor([I] * 3, [0, 1, 2]), sr.into_executor([I] * 3, [0, 1, 2]),
                            JoinParams([0, 1, 2], [0, 1, 2]), JoinParams([0, 1, 2], [0, 1, 2]), [False, False, False])
    mv = Counter()

    def ins(rows, ops=None):
        cols = list(zip(*rows))
        return StreamChunk(np.full(len(rows), abi.OP_INSERT, np.uint8) if ops is None else ops,
                           [Column(I, np.array(c, dtype=np.int64)) for c in cols])

    def sink(chunks):
        for ch in chunks:
            for op, row in ch.rows():
                mv[(row[0], row[1])] += 1 if op in (abi.OP_INSERT, abi.OP_UPDATE_INSERT) else -1

    def window(t):
        return (t // W) * W, (t // W) * W + W

    person = [[r[0], *window(r[2])] for r in fx["person"]]
    auction = [[r[0], *window(r[1])] for r in fx["auction"]]
    for e in range(5):
        for side, rows in ((0, person[e * 4:(e + 1) * 4]), (1, auction[e * 8:(e + 1) * 8])):
            if rows:
                aggs[side].apply_chunk(ins(rows))
            for ch in aggs[side].flush_data(e + 1):  # (k0, k1, k2, count): the count is projected away
                keyrows = [list(row[:3]) for _, row in ch.rows()]
                ops = np.array([op for op, _ in ch.rows()], dtype=np.uint8)
                if keyrows:
                    sink(join.eq_join_oneside(side, ins(keyrows, ops)))
    name = {r[0]: r[1] for r in fx["person"]}
    assert all(v in (0, 1) for v in mv.values())
    got = sorted([k[0], name[k[0]], k[1]] for k, v in mv.items() if v)
    assert got == sorted(fx["expected_q8"])


def run_tpch_q3(backend, n_cust=300, n_orders=1500, n_items=6000, seed=0x7C43, epochs=6):
    """SURVEY 8(d) cfg5, the TPC-H q3 streaming plan (src/frontend/planner_test/tests/testdata/output/tpch.yaml,
    `tpch_q3` stream_plan) at test size, incrementally through the operators of `backend`:
        customer [c_mktsegment filter upstream] JOIN orders ON c_custkey = o_custkey  [o_orderdate < D1]
          -> Project(o_orderkey, o_orderdate, o_shippriority)
          JOIN lineitem ON l_orderkey = o_orderkey                                    [l_shipdate > D1]
          -> Project(l_orderkey, o_orderdate, o_shippriority, l_extendedprice * (100 - l_discount))
          -> HashAgg(sum, count GROUP BY l_orderkey, o_orderdate, o_shippriority)      (Key128-class key: i64, date, i32)
    Money is scale-2 fixed point in int64 (revenue = scale 4), the sum accumulates in 128 bits and leaves as a decimal.
    The streams carry retractions (orders are deleted again, line items updated), so both joins emit Delete rows and
    the aggregation retracts.  Returns (materialized view {key: (count, revenue)}, per-barrier delta multisets); the view
    is checked here against a direct evaluation of the SQL over the rows that are live at the end."""
    from collections import Counter
    from risingwave_b200.executor import AggCall, FilterExecutor, HashAggExecutor, HashJoinExecutor, JoinParams, MockSource, ProjectExecutor
    from risingwave_b200.stream_chunk import Column, StreamChunk
    I, D, I4 = abi.T_INT64, abi.T_DATE, abi.T_INT32
    NPT = {I: np.int64, D: np.int32, I4: np.int32}
    rng = np.random.default_rng(seed)
    D1 = 9204  # 1995-03-15 as days since 1970-01-01
    # ---- tables (customer: only the rows that pass the segment filter reach the join)
    cust = [(int(k),) for k in rng.permutation(n_cust * 5)[:n_cust]]
    cust_keys = np.array([c[0] for c in cust] + list(range(n_cust * 5, n_cust * 5 + 40)))  # some orders of filtered-out customers
    orders = [(int(ok), int(rng.choice(cust_keys)), int(D1 + rng.integers(-60, 20)), int(rng.integers(0, 3)))
              for ok in rng.permutation(n_orders * 4)[:n_orders]]
    okeys = np.array([o[0] for o in orders])
    items = []
    seen = Counter()
    for _ in range(n_items):
        ok = int(rng.choice(okeys))
        seen[ok] += 1
        items.append((ok, int(rng.integers(100, 10_000_000)), int(rng.integers(0, 11)), int(D1 + rng.integers(-20, 60)), seen[ok]))
    T_ORD, T_ITEM = [I, I, D, I4], [I, I, I, D, I]

    def chunk(rows, types, ops=None):
        cols = list(zip(*rows))
        return StreamChunk(np.full(len(rows), abi.OP_INSERT, np.uint8) if ops is None else np.asarray(ops, np.uint8),
                           [Column(t, np.array(c, dtype=NPT[t])) for t, c in zip(types, cols)])

    def src(types, pk):
        _, s = MockSource.channel()
        return s.into_executor(types, pk)

    f_ord = FilterExecutor(backend, src(T_ORD, [0]), f"(less_than:boolean $2:date {D1}:date)")
    f_item = FilterExecutor(backend, src(T_ITEM, [0, 4]), f"(greater_than:boolean $3:date {D1}:date)")
    # join 1: customer (key 0) x orders (key col 1 = o_custkey, stream key o_orderkey)
    j1 = HashJoinExecutor(backend, abi.JOIN_INNER, src([I], [0]), src(T_ORD, [0]), JoinParams([0], [0]), JoinParams([1], [0]), [False])
    p1 = ProjectExecutor(backend, src([I] + T_ORD, []), ["$1:int8", "$3:date", "$4:int4"])
    # join 2: (o_orderkey, o_orderdate, o_shippriority) x lineitem (key col 0, stream key (l_orderkey, l_linenumber))
    j2 = HashJoinExecutor(backend, abi.JOIN_INNER, src([I, D, I4], [0]), src(T_ITEM, [0, 4]), JoinParams([0], [0]), JoinParams([0], [0, 4]), [False])
    p2 = ProjectExecutor(backend, src([I, D, I4] + T_ITEM, []),
                         ["$3:int8", "$1:date", "$2:int4", "(multiply:int8 $4:int8 (subtract:int8 100:int8 $5:int8))"])
    agg = HashAggExecutor(backend, src([I, D, I4, I], []), False,
                          [AggCall.from_pretty(c) for c in ("(count:int8)", "(sum:decimal $3:int8)")], 0, [0, 1, 2])
    mv, deltas = {}, []

    def to_agg(chunks):
        for ch in chunks:
            if ch.cardinality():
                agg.apply_chunk(p2.apply_project_exprs(ch))

    def push_orders(rows, ops=None):
        f = f_ord.filter(chunk(rows, T_ORD, ops))
        if f is None:
            return
        for ch in j1.eq_join_oneside(1, f):
            if ch.cardinality():
                to_agg(j2.eq_join_oneside(0, p1.apply_project_exprs(ch)))

    def push_items(rows, ops=None):
        f = f_item.filter(chunk(rows, T_ITEM, ops))
        if f is not None:
            to_agg(j2.eq_join_oneside(1, f))

    def push_customers(rows):
        for ch in j1.eq_join_oneside(0, chunk(rows, [I])):
            if ch.cardinality():
                to_agg(j2.eq_join_oneside(0, p1.apply_project_exprs(ch)))

    def barrier(epoch):
        d = Counter()
        for ch in agg.flush_data(epoch):
            for op, row in ch.rows():
                key, val = tuple(row[:3]), (row[3], row[4])
                d[(op, key, val)] += 1
                if op in (abi.OP_INSERT, abi.OP_UPDATE_INSERT):
                    mv[key] = val
                else:
                    assert mv.get(key) == val, (key, val, mv.get(key))
                    del mv[key]
        deltas.append(d)

    live_orders, live_items = {}, {}
    co = len(cust) // epochs + 1
    oo = len(orders) // epochs + 1
    io = len(items) // epochs + 1
    for e in range(epochs):
        push_customers(cust[e * co:(e + 1) * co])
        new_o = orders[e * oo:(e + 1) * oo]
        if new_o:
            push_orders(new_o)
            live_orders.update({o[0]: o for o in new_o})
        new_i = items[e * io:(e + 1) * io]
        if new_i:
            push_items(new_i)
            live_items.update({(i[0], i[4]): i for i in new_i})
        # retractions: a few orders disappear, a few line items change their discount (U- / U+)
        gone = [live_orders.pop(k) for k in list(live_orders)[:: 17][:20]]
        if gone:
            push_orders(gone, [abi.OP_DELETE] * len(gone))
        upd = [live_items[k] for k in list(live_items)[:: 23][:30]]
        if upd:
            rows, ops = [], []
            for it in upd:
                new = (it[0], it[1], (it[2] + 3) % 11, it[3], it[4])
                rows += [it, new]
                ops += [abi.OP_UPDATE_DELETE, abi.OP_UPDATE_INSERT]
                live_items[(it[0], it[4])] = new
            push_items(rows, ops)
        barrier(e + 1)
    # ---- the SQL, evaluated directly over the live rows
    custset = {c[0] for c in cust}
    want = {}
    for it in live_items.values():
        o = live_orders.get(it[0])
        if o is None or it[3] <= D1 or o[2] >= D1 or o[1] not in custset:
            continue
        key = (o[0], o[2], o[3])
        n, s = want.get(key, (0, 0))
        want[key] = (n + 1, s + it[1] * (100 - it[2]))
    got = {k: (v[0], int(v[1])) for k, v in mv.items()}
    assert got == want, (len(got), len(want))
    return got, deltas
