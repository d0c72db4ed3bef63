"""Host-side mirror of the reference's StreamChunk data model.

StreamChunk{ops, DataChunk{columns, visibility}}  src/common/src/array/stream_chunk.rs:106-110,
data_chunk.rs:65-68; Op stream_chunk.rs:45-50; the `from_pretty` test DSL stream_chunk.rs:650-750 /
data_chunk.rs:708-790 (so the reference's golden test literals can be used verbatim).

Columns are numpy arrays in the ABI's native widths; validity / visibility are bool arrays here and
are bit-packed LSB-first (bitmap.rs:363-369) only when crossing the C ABI.
"""
from __future__ import annotations

import ctypes as C
from collections import Counter
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import abi

DECIMAL_DTYPE = np.dtype([("lo", "<u8"), ("hi", "<i8")])

NP_DTYPE = {
    abi.T_BOOL: np.dtype(np.uint8), abi.T_INT16: np.dtype(np.int16), abi.T_INT32: np.dtype(np.int32),
    abi.T_INT64: np.dtype(np.int64), abi.T_FLOAT32: np.dtype(np.float32),
    abi.T_FLOAT64: np.dtype(np.float64), abi.T_DATE: np.dtype(np.int32), abi.T_TIME: np.dtype(np.int64),
    abi.T_TIMESTAMP: np.dtype(np.int64), abi.T_TIMESTAMPTZ: np.dtype(np.int64),
    abi.T_SERIAL: np.dtype(np.int64), abi.T_DECIMAL: DECIMAL_DTYPE,
}

# from_pretty type tokens (data_chunk.rs:726-741); "s" (int16) is our extension
PRETTY_TYPES = {"B": abi.T_BOOL, "s": abi.T_INT16, "i": abi.T_INT32, "I": abi.T_INT64,
                "f": abi.T_FLOAT32, "F": abi.T_FLOAT64, "D": abi.T_DATE, "TS": abi.T_TIMESTAMP,
                "TZ": abi.T_TIMESTAMPTZ, "SRL": abi.T_SERIAL, "DEC": abi.T_DECIMAL, "T": abi.T_VARCHAR}
# varlen columns (BytesArray, bytes_array.rs:30-34) hold Python `bytes` in an object array on the host
for _t in abi.VARLEN_TYPES:
    NP_DTYPE[_t] = np.dtype(object)
PRETTY_TOKENS = {v: k for k, v in PRETTY_TYPES.items()}
OP_TOKENS = {"+": abi.OP_INSERT, "-": abi.OP_DELETE, "U+": abi.OP_UPDATE_INSERT, "U-": abi.OP_UPDATE_DELETE}
OP_STR = {v: k for k, v in OP_TOKENS.items()}


def pack_bits(b: np.ndarray) -> np.ndarray:
    """bool[n] -> uint64 words, LSB first, zero padded."""
    n = len(b)
    nwords = max(1, (n + 63) // 64)
    by = np.packbits(np.asarray(b, dtype=np.uint8), bitorder="little")
    out = np.zeros(nwords * 8, dtype=np.uint8)
    out[: len(by)] = by
    return out.view(np.uint64)


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n].astype(bool)


def decimal_to_int(v) -> int:
    lo, hi = int(v["lo"]), int(v["hi"])
    return (hi << 64) | lo


def int_to_decimal(x: int):
    lo = x & ((1 << 64) - 1)
    hi = x >> 64
    return (lo, hi)


@dataclass
class Column:
    type: int
    data: np.ndarray
    valid: Optional[np.ndarray] = None  # bool[n]; None = no NULLs

    def value(self, i: int):
        if self.valid is not None and not self.valid[i]:
            return None
        v = self.data[i]
        if self.type in abi.VARLEN_TYPES:
            return bytes(v)
        if self.type == abi.T_DECIMAL:
            return decimal_to_int(v)
        if self.type in (abi.T_FLOAT32, abi.T_FLOAT64):
            return float(v)
        if self.type == abi.T_BOOL:
            return bool(v)
        return int(v)


class StreamChunk:
    def __init__(self, ops, columns: Sequence[Column], vis: Optional[np.ndarray] = None):
        self.ops = np.ascontiguousarray(ops, dtype=np.uint8)
        self.columns: List[Column] = list(columns)
        self.vis = None if vis is None else np.ascontiguousarray(vis, dtype=bool)
        n = len(self.ops)
        for c in self.columns:
            assert len(c.data) == n, "column length mismatch"

    # ------------------------------------------------------------------ basics
    def capacity(self) -> int:
        return len(self.ops)

    def cardinality(self) -> int:
        return len(self.ops) if self.vis is None else int(self.vis.sum())

    def types(self) -> List[int]:
        return [c.type for c in self.columns]

    def is_visible(self, i: int) -> bool:
        return True if self.vis is None else bool(self.vis[i])

    def row(self, i: int) -> Tuple:
        return tuple(c.value(i) for c in self.columns)

    def rows(self, include_invisible: bool = False):
        """(op, row) for each visible row (StreamChunk::rows)."""
        for i in range(len(self.ops)):
            if include_invisible or self.is_visible(i):
                yield int(self.ops[i]), self.row(i)

    # ------------------------------------------------------------------ from_pretty / to_pretty
    @staticmethod
    def from_pretty(s: str) -> "StreamChunk":
        lines = [ln.strip() for ln in s.split("\n") if ln.strip()]
        header = []
        for tok in lines[0].split():
            if tok == "//":
                break
            header.append(PRETTY_TYPES[tok])
        ops, vis = [], []
        vals = [[] for _ in header]
        for ln in lines[1:]:
            toks = ln.split()
            if toks[0] == "//":
                continue
            ops.append(OP_TOKENS[toks[0]])
            rest = toks[1:]
            for k, t in enumerate(header):
                vals[k].append(_parse_value(rest[k], t))
            tail = rest[len(header):]
            if not tail or tail[0] == "//":
                vis.append(True)
            elif tail[0] == "D":
                vis.append(False)
            else:
                raise ValueError(f"invalid token {tail[0]!r}")
        cols = [column_from_values(t, v) for t, v in zip(header, vals)]
        v = np.array(vis, dtype=bool)
        return StreamChunk(np.array(ops, dtype=np.uint8), cols, None if v.all() else v)

    def to_pretty(self) -> str:
        out = [" ".join(PRETTY_TOKENS[c.type] for c in self.columns)]
        for i in range(len(self.ops)):
            vals = ["." if v is None else _fmt(v) for v in self.row(i)]
            out.append(" ".join([OP_STR[int(self.ops[i])]] + vals + ([] if self.is_visible(i) else ["D"])))
        return "\n".join(out)

    __repr__ = to_pretty

    # ------------------------------------------------------------------ equality as in the reference
    # (derive(PartialEq): ops, all column values AND visibility; stream_chunk.rs:104-110)
    def __eq__(self, other) -> bool:
        if not isinstance(other, StreamChunk):
            return NotImplemented
        if self.types() != other.types() or len(self.ops) != len(other.ops):
            return False
        if not np.array_equal(self.ops, other.ops):
            return False
        for i in range(len(self.ops)):
            if self.is_visible(i) != other.is_visible(i):
                return False
            if not _row_eq(self.row(i), other.row(i)):
                return False
        return True

    def sort_rows(self) -> "StreamChunk":
        """visible rows sorted by (op, row) like snapshot.rs `sort_chunk` (compacts the chunk)."""
        rs = sorted(self.rows(), key=lambda r: (r[0], tuple((v is None, 0 if v is None else v) for v in r[1])))
        return StreamChunk.from_rows(self.types(), rs)

    @staticmethod
    def from_rows(types: Sequence[int], rows: Sequence[Tuple[int, Tuple]]) -> "StreamChunk":
        ops = np.array([r[0] for r in rows], dtype=np.uint8)
        cols = [column_from_values(t, [r[1][k] for r in rows]) for k, t in enumerate(types)]
        return StreamChunk(ops, cols)

    def slice(self, lo: int, hi: int) -> "StreamChunk":
        cols = [Column(c.type, c.data[lo:hi], None if c.valid is None else c.valid[lo:hi]) for c in self.columns]
        return StreamChunk(self.ops[lo:hi], cols, None if self.vis is None else self.vis[lo:hi])

    # ------------------------------------------------------------------ C ABI
    def to_abi(self):
        """-> (RwChunk, keepalive). Buffers are borrowed by the callee for the duration of a call."""
        keep = []
        n = len(self.ops)
        cols = (abi.RwColumn * max(1, len(self.columns)))()
        for k, c in enumerate(self.columns):
            cols[k].type = c.type
            if c.type in abi.VARLEN_TYPES:
                vals = [b"" if (c.valid is not None and not c.valid[i]) else bytes(c.data[i]) for i in range(n)]
                offs = np.zeros(n + 1, dtype=np.uint32)
                np.cumsum([len(v) for v in vals], out=offs[1:])
                blob = np.frombuffer(b"".join(vals) + b"\0", dtype=np.uint8).copy()
                keep += [offs, blob]
                cols[k].data = blob.ctypes.data
                cols[k].offsets = offs.ctypes.data
            else:
                data = np.ascontiguousarray(c.data, dtype=NP_DTYPE[c.type])
                keep.append(data)
                cols[k].data = data.ctypes.data if n else None
            if c.valid is not None and not bool(np.all(c.valid)):
                w = pack_bits(c.valid)
                keep.append(w)
                cols[k].validity = w.ctypes.data
            else:
                cols[k].validity = None
        ch = abi.RwChunk()
        ch.n_rows = n
        ch.n_cols = len(self.columns)
        ops = self.ops
        keep.append(ops)
        ch.ops = ops.ctypes.data if n else None
        if self.vis is not None and not bool(np.all(self.vis)):
            w = pack_bits(self.vis)
            keep.append(w)
            ch.visibility = w.ctypes.data
        else:
            ch.visibility = None
        ch.columns = cols
        keep.append(cols)
        return ch, keep

    @staticmethod
    def from_abi(view: abi.RwChunk) -> "StreamChunk":
        """Copy a host rw_chunk view into an owned StreamChunk."""
        n = int(view.n_rows)
        ops = np.ctypeslib.as_array(C.cast(view.ops, C.POINTER(C.c_uint8)), shape=(n,)).copy() if n else np.zeros(0, np.uint8)
        nw = max(1, (n + 63) // 64)
        vis = None
        if view.visibility:
            w = np.ctypeslib.as_array(C.cast(view.visibility, C.POINTER(C.c_uint64)), shape=(nw,)).copy()
            vis = unpack_bits(w, n)
        cols = []
        for k in range(view.n_cols):
            c = view.columns[k]
            dt = NP_DTYPE[c.type]
            if c.type in abi.VARLEN_TYPES:
                data = np.empty(n, dtype=object)
                if n:
                    offs = np.ctypeslib.as_array(C.cast(c.offsets, C.POINTER(C.c_uint32)), shape=(n + 1,)).copy()
                    total = int(offs[n])
                    blob = bytes(np.ctypeslib.as_array(C.cast(c.data, C.POINTER(C.c_uint8)), shape=(max(total, 1),))[:total])
                    for i in range(n):
                        data[i] = blob[int(offs[i]):int(offs[i + 1])]
            elif n:
                raw = np.ctypeslib.as_array(C.cast(c.data, C.POINTER(C.c_uint8)), shape=(n * dt.itemsize,)).copy()
                data = raw.view(dt)
            else:
                data = np.zeros(0, dt)
            valid = None
            if c.validity:
                w = np.ctypeslib.as_array(C.cast(c.validity, C.POINTER(C.c_uint64)), shape=(nw,)).copy()
                valid = unpack_bits(w, n)
            cols.append(Column(int(c.type), data, valid))
        return StreamChunk(ops, cols, vis)


# ---------------------------------------------------------------------- helpers
def _parse_value(tok: str, t: int):
    if tok == ".":
        return None
    if t in (abi.T_FLOAT32, abi.T_FLOAT64):
        return float(tok)
    if t == abi.T_BOOL:
        return tok in ("t", "true", "1", "T")
    if t in abi.VARLEN_TYPES:
        return b"" if tok == "(empty)" else tok.encode()
    return int(tok)


def _fmt(v) -> str:
    if isinstance(v, bool):
        return "t" if v else "f"
    if isinstance(v, float):
        return repr(v)
    if isinstance(v, bytes):
        return v.decode(errors="replace") if v else "(empty)"
    return str(v)


def _row_eq(a: Tuple, b: Tuple) -> bool:
    for x, y in zip(a, b):
        if x is None or y is None:
            if x is not y:
                return False
        elif isinstance(x, float) and isinstance(y, float) and x != x and y != y:
            continue
        elif x != y:
            return False
    return True


def column_from_values(t: int, vals: Sequence) -> Column:
    n = len(vals)
    dt = NP_DTYPE[t]
    data = np.zeros(n, dtype=dt)
    valid = np.ones(n, dtype=bool)
    for i, v in enumerate(vals):
        if v is None:
            valid[i] = False
            if t in abi.VARLEN_TYPES:
                data[i] = b""
        elif t in abi.VARLEN_TYPES:
            data[i] = v.encode() if isinstance(v, str) else bytes(v)
        elif t == abi.T_DECIMAL:
            data[i] = int_to_decimal(int(v))
        else:
            data[i] = v
    return Column(t, data, None if valid.all() else valid)


def column_from_numpy(t: int, arr: np.ndarray, valid: Optional[np.ndarray] = None) -> Column:
    return Column(t, np.ascontiguousarray(arr, dtype=NP_DTYPE[t]), valid)


def concat_chunks(chunks: Sequence[StreamChunk]) -> StreamChunk:
    assert chunks
    types = chunks[0].types()
    ops = np.concatenate([c.ops for c in chunks])
    any_vis = any(c.vis is not None for c in chunks)
    vis = np.concatenate([c.vis if c.vis is not None else np.ones(len(c.ops), bool) for c in chunks]) if any_vis else None
    cols = []
    for k, t in enumerate(types):
        data = np.concatenate([c.columns[k].data for c in chunks])
        any_v = any(c.columns[k].valid is not None for c in chunks)
        valid = np.concatenate([c.columns[k].valid if c.columns[k].valid is not None else np.ones(len(c.ops), bool)
                                for c in chunks]) if any_v else None
        cols.append(Column(t, data, valid))
    return StreamChunk(ops, cols, vis)


def net_multiset(chunks: Sequence[StreamChunk]) -> Counter:
    """Net applied change of a sequence of chunks: Counter[row] = (#inserts - #deletes) over visible
    rows, zero entries dropped.  This is the `Store::apply_chunk` comparator of the reference's
    snapshot tests (src/stream/tests/integration_tests/snapshot.rs:219-254)."""
    c: Counter = Counter()
    for ch in chunks:
        for op, row in ch.rows():
            key = tuple(("nan" if isinstance(v, float) and v != v else v) for v in row)
            c[key] += 1 if op in (abi.OP_INSERT, abi.OP_UPDATE_INSERT) else -1
    return Counter({k: v for k, v in c.items() if v != 0})


def emitted_multiset(chunks: Sequence[StreamChunk]) -> Counter:
    """Multiset of visible (op-class, row) actually emitted (Insert/UpdateInsert vs Delete/UpdateDelete
    are kept distinct from each other but U+/+ are identified, as downstream semantics do)."""
    c: Counter = Counter()
    for ch in chunks:
        for op, row in ch.rows():
            key = tuple(("nan" if isinstance(v, float) and v != v else v) for v in row)
            c[(op in (abi.OP_INSERT, abi.OP_UPDATE_INSERT), key)] += 1
    return c
