"""GPU debugging aid (not a test): the bench's retract leg at a configurable scale, with the output of every update push
compared ROW BY ROW against a numpy restatement (for each update row: all stored bids of its auction)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from risingwave_b200 import abi, device  # noqa: E402
from risingwave_b200.executor import Backend, HashJoinExecutor, JoinParams, MockSource  # noqa: E402

N_BUILD = int(os.environ.get("DBG_BUILD", 1 << 20))
BATCH = int(os.environ.get("DBG_BATCH", 1 << 18))
N_BATCH = int(os.environ.get("DBG_NBATCH", 20))
RP = int(os.environ.get("DBG_PAIRS", 1 << 17))
STEPS = int(os.environ.get("DBG_STEPS", 3))
T4 = [abi.T_INT64] * 4
be = Backend.cuda()
stream = torch.cuda.Stream()


def to_dev(cols):
    return [torch.from_numpy(c).cuda() for c in cols]


with torch.cuda.stream(stream):
    _, sl = MockSource.channel()
    _, sr = MockSource.channel()
    join = HashJoinExecutor(be, abi.JOIN_INNER, sl.into_executor(T4, [1]), sr.into_executor(T4, [0]), JoinParams([0], [1]), JoinParams([0], []),
                            [False], capacity_hint=(N_BUILD, N_BUILD))
    auct = bench.gen_auctions(N_BUILD, bench.SEED)
    ad = to_dev(auct)
    device.join_push_device(join, abi.SIDE_RIGHT, device.DeviceChunk(torch.ones(N_BUILD, dtype=torch.uint8, device="cuda"), ad, T4), stream)
    bids = [bench.gen_bids(BATCH, s * BATCH, bench.SEED, N_BUILD) for s in range(N_BATCH)]
    PIPE = int(os.environ.get("DBG_PIPE", 0))  # 1: two pushes outstanding, as bench.py drives the handle
    if PIPE:
        bdev = [device.DeviceChunk(torch.ones(BATCH, dtype=torch.uint8, device="cuda"), to_dev(b), T4) for b in bids]
        for i, ch in enumerate(bdev):
            device.join_push_device_async(join, abi.SIDE_LEFT, ch, stream)
            if i:
                assert device.join_collect(join, stream).n_rows == BATCH
        assert device.join_collect(join, stream).n_rows == BATCH
    else:
        for b in bids:
            o = device.join_push_device(join, abi.SIDE_LEFT, device.DeviceChunk(torch.ones(BATCH, dtype=torch.uint8, device="cuda"), to_dev(b), T4), stream)
            assert o.n_rows == BATCH, o.n_rows
    allb = [np.concatenate([b[k] for b in bids]) for k in range(4)]
    order = np.argsort(allb[0], kind="stable")
    sk = allb[0][order]
    bad = 0
    ups = [bench.gen_auction_updates(auct, s * RP, RP) for s in range(STEPS)]
    udev = [device.DeviceChunk(torch.from_numpy(o_).cuda(), to_dev(c_), T4) for o_, c_ in ups]
    outs = {}
    cks = []
    if PIPE:  # every step but the last with two outstanding; outputs snapshotted at collect
        def snap(o):
            torch.cuda.synchronize()
            vis = o.visible()
            cks.append(o.checksum(bench.CHECKSUM_WEIGHTS))
            return (o.n_rows, o.ops().cpu().numpy(), np.stack([o.column(k).cpu().numpy() for k in range(8)], 1), None if vis is None else vis.cpu().numpy())
        for s in range(STEPS - 1):
            device.join_push_device_async(join, abi.SIDE_RIGHT, udev[s], stream)
            if s:
                outs[s - 1] = snap(device.join_collect(join, stream))
        if STEPS > 1:
            outs[STEPS - 2] = snap(device.join_collect(join, stream))
        device.join_push_device_async(join, abi.SIDE_RIGHT, udev[STEPS - 1], stream)
        outs[STEPS - 1] = snap(device.join_collect(join, stream))
    for s in range(STEPS):
        ops, cols = ups[s]
        if PIPE:
            n_out, got_ops, got, v = outs[s]
            if v is not None:
                got_ops, got = got_ops[v], got[v]

            class _O:
                n_rows = n_out
            o = _O()
        else:
            ch = udev[s]
            if s % 2 == 0:
                o = device.join_push_device(join, abi.SIDE_RIGHT, ch, stream)
            else:
                device.join_push_device_async(join, abi.SIDE_RIGHT, ch, stream)
                o = device.join_collect(join, stream)
            torch.cuda.synchronize()
            vis = o.visible()
            got_ops = o.ops().cpu().numpy()
            got = np.stack([o.column(k).cpu().numpy() for k in range(8)], 1)
            if vis is not None:
                v = vis.cpu().numpy()
                got_ops, got = got_ops[v], got[v]
        # expectation
        lo = np.searchsorted(sk, cols[0], "left")
        hi = np.searchsorted(sk, cols[0], "right")
        cnt = hi - lo
        rep = np.repeat(np.arange(len(ops)), cnt)
        off = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        bi = order[np.repeat(lo, cnt) + off]
        want = np.stack([allb[k][bi] for k in range(4)] + [cols[k][rep] for k in range(4)], 1)
        want_ops = np.where(ops[rep] == 4, 2, 1).astype(np.uint8)
        g = np.concatenate([got_ops[:, None].astype(np.int64), got], 1)
        w = np.concatenate([want_ops[:, None].astype(np.int64), want], 1)
        gs = g[np.lexsort(g.T[::-1])]
        ws = w[np.lexsort(w.T[::-1])]
        same = gs.shape == ws.shape and bool((gs == ws).all())
        wsum = int((np.where(w[:, 0] == 1, 1, -1).astype(np.int64)[:, None] * w[:, 1:] * np.array(bench.CHECKSUM_WEIGHTS, np.int64)[None, :]).sum()) & ((1 << 64) - 1)
        print(f"step {s}: out {o.n_rows} visible {len(g)} expected {len(w)} equal {same}  expected checksum {wsum:016x}"
              + (f"  DeviceView.checksum {cks[s][0]} {cks[s][1]:016x}" if PIPE else ""), flush=True)
        if not same:
            bad += 1
            if gs.shape == ws.shape:
                d = np.nonzero((gs != ws).any(1))[0]
                print("  differing rows:", len(d), "first:", d[:5])
                for i in d[:5]:
                    print("   got ", gs[i].tolist())
                    print("   want", ws[i].tolist())
            gu, gc = np.unique(g[:, 0], return_counts=True)
            wu, wc = np.unique(w[:, 0], return_counts=True)
            print("  ops got", dict(zip(gu.tolist(), gc.tolist())), "want", dict(zip(wu.tolist(), wc.tolist())))
            for k in range(1, 9):
                print(f"  col {k - 1}: sum got {int(g[:, k].sum())} want {int(w[:, k].sum())}  set-equal {bool(np.array_equal(np.sort(g[:, k]), np.sort(w[:, k])))}")
print("BAD" if bad else "ALL EQUAL")
