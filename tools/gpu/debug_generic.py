"""GPU debugging aid: bench.py's generic (LEFT OUTER) leg at bench scale, the last push compared with numpy column by column,
per-push wall clock printed."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from risingwave_b200 import abi, device  # noqa: E402
from risingwave_b200.executor import Backend, HashJoinExecutor, JoinParams, MockSource  # noqa: E402

NG, BG, NP = int(os.environ.get("DBG_NG", 1 << 20)), int(os.environ.get("DBG_BG", 1 << 18)), int(os.environ.get("DBG_NP", 10))
T4 = [abi.T_INT64] * 4
be = Backend.cuda()
stream = torch.cuda.Stream()


def dchunk(cols):
    return device.DeviceChunk(torch.ones(len(cols[0]), dtype=torch.uint8, device="cuda"), [torch.from_numpy(c).cuda() for c in cols], T4)


with torch.cuda.stream(stream):
    _, gl = MockSource.channel()
    _, gr = MockSource.channel()
    jg = HashJoinExecutor(be, abi.JOIN_LEFT_OUTER, gl.into_executor(T4, [1]), gr.into_executor(T4, [0]), JoinParams([0], [1]), JoinParams([0], []),
                          [False], capacity_hint=(NG, NG))
    ag = bench.gen_auctions(NG, bench.SEED + 77)
    t0 = time.perf_counter()
    device.join_push_device(jg, abi.SIDE_RIGHT, dchunk(ag), stream)
    torch.cuda.synchronize()
    print(f"build push {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
    pos_of = np.empty(NG, np.int64)
    pos_of[ag[0]] = np.arange(NG)
    for s in range(NP):
        vb = bench.gen_bids(BG, s * BG, bench.SEED + 5, NG + NG // 8)
        ch = dchunk(vb)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vv = device.join_push_device(jg, abi.SIDE_LEFT, ch, stream)
        torch.cuda.synchronize()
        dt = 1e3 * (time.perf_counter() - t0)
        got = [vv.column(k).cpu().numpy() for k in range(8)]
        ops = vv.ops().cpu().numpy()
        vis = vv.visible()
        valid = []
        for k in range(8):
            if vv.valid_ptrs[k]:
                nw = (vv.n_rows + 63) // 64
                w = torch.empty(nw, dtype=torch.int64, device="cuda")
                device._d2d(w.data_ptr(), vv.valid_ptrs[k], nw * 8)
                valid.append(np.unpackbits(w.cpu().numpy().view(np.uint8), bitorder="little")[:vv.n_rows].astype(bool))
            else:
                valid.append(None)
        matched = vb[0] < NG
        msg = [f"push {s}: {dt:.2f} ms, out {vv.n_rows}, vis {'none' if vis is None else int(vis.sum().item())}, ops {np.unique(ops).tolist()}"]
        if vv.n_rows == BG:
            og, ow = np.argsort(got[1], kind="stable"), np.argsort(vb[1], kind="stable")
            for k in range(4):
                msg.append(f"c{k}:{bool(np.array_equal(got[k][og], vb[k][ow]))}")
            mw = matched[ow]
            for k in range(4):
                want_k = ag[k][pos_of[np.where(mw, vb[0][ow], 0)]]
                eq = bool(np.array_equal(got[4 + k][og][mw], want_k[mw]))
                nv = "no-validity" if valid[4 + k] is None else f"valid==matched:{bool(np.array_equal(valid[4 + k][og], mw))}"
                msg.append(f"c{4 + k}:{eq},{nv}")
                if not eq:
                    bad = np.nonzero(got[4 + k][og][mw] != want_k[mw])[0]
                    msg.append(f"(bad {len(bad)} first got {got[4 + k][og][mw][bad[:3]].tolist()} want {want_k[mw][bad[:3]].tolist()})")
        print(" ".join(msg), flush=True)
    new_ids = NG + np.arange(0, NG // 8, dtype=np.int64)
    upd = [new_ids, new_ids % 1000, 10 + new_ids % 5, new_ids * 3]
    ch = dchunk(upd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vu = device.join_push_device(jg, abi.SIDE_RIGHT, ch, stream)
    torch.cuda.synchronize()
    print(f"flip push {1e3 * (time.perf_counter() - t0):.2f} ms, out {vu.n_rows}", flush=True)
