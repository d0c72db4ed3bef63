// join.cu -- streaming two-sided incremental HashJoin on sm_100a.
//
// Replaces (reference, Rust):
//   HashJoinExecutor::eq_join_oneside      src/stream/src/executor/hash_join.rs:925-1062
//   handle_match_rows / handle_match_row   hash_join.rs:1072-1357
//   JoinChunkBuilder (output-op rules)     src/stream/src/executor/join/builder.rs:158-316
//   JoinHashMap / JoinEntryState           src/stream/src/executor/join/hash_join.rs:169-830
//   join-type predicates                   src/stream/src/executor/join/mod.rs:103-169
//
// HBM layout per side:
//   row store  : columnar, append-only; col[c][row] in native width, optional valid byte / row,
//                link[row] = next row of the same key | DEAD bit, degree[row] (u32) when needed
//   hash index : open addressing, linear probing, power-of-two capacity, load <= 1/2;
//                slot = key word(s) | (count << 32 | head)       (Key64: 16 B)
// Two execution paths:
//   * inner fast path (no degrees): row-parallel probe; matches are emitted with block-scan
//     compaction; own-side inserts / deletes are applied by separate row-parallel kernels.
//   * generic path (all 8 join types, degrees, append-only optimisation, mixed +/- on one key):
//     the batch is grouped by join key (scratch hash table + radix sort) and ONE thread walks each
//     key's rows in input order -- state of different keys is disjoint, so this is exactly the
//     reference's sequential semantics with the parallelism taken across keys.
#include <algorithm>
#include <memory>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace rw {

#define J_EMPTY 0x8000000000000000ull
#define J_NIL 0x7fffffffu
#define J_DEAD 0x80000000u
#define J_MAX_OUT (2 * RW_MAX_COLS)

#define JERR_DOUBLE_DELETE 1u
#define JERR_OUT_CAPACITY 2u
#define JERR_APPEND_ONLY_MULTI 4u

struct JoinPlanDev {
  int T;
  int n_keys;
  int key_col[2][RW_MAX_KEYS];
  int null_safe[RW_MAX_KEYS];
  int n_cols[2];
  int col_type[2][RW_MAX_COLS];
  int col_width[2][RW_MAX_COLS];
  int n_pk[2];
  int pk_col[2][RW_MAX_COLS];
  int n_out;
  int out_type[J_MAX_OUT];
  int out_width[J_MAX_OUT];
  int n_map[2];
  int map_in[2][J_MAX_OUT];
  int map_out[2][J_MAX_OUT];
  int need_degree[2];
  int append_only_optimize;
  int cond_cmp, cond_lhs, cond_rhs;
  int single_key, KW, SW;
  int strict;
};

struct JoinSideDev {
  void* col[RW_MAX_COLS];
  uint8_t* valid[RW_MAX_COLS];  // nullptr until the column has seen a NULL
  uint32_t* link;
  uint32_t* degree;  // nullptr if the side keeps no degrees
  uint64_t* slots;
  uint64_t cap;
};

struct JoinStatus {
  unsigned long long out_rows;    // rows reserved in the output
  unsigned long long n_store;     // store candidates of this push
  unsigned long long n_keys[2];   // distinct keys ever claimed per side
  unsigned long long live_rows[2];
  unsigned int err;
  unsigned int pad;
};

struct JoinOutDev {
  uint8_t* ops;
  uint8_t* vis;
  void* col[J_MAX_OUT];
  uint8_t* valid[J_MAX_OUT];
  unsigned int* has_null;  // [J_MAX_OUT] + [J_MAX_OUT] = any invisible flag
  int64_t capacity;
};

// join/mod.rs:103-169
__device__ __host__ __forceinline__ bool jt_is_outer_side(int T, int S) { return T == RW_JOIN_FULL_OUTER || (T == RW_JOIN_LEFT_OUTER && S == 0) || (T == RW_JOIN_RIGHT_OUTER && S == 1); }
__device__ __host__ __forceinline__ bool jt_outer_side_null(int T, int S) { return T == RW_JOIN_FULL_OUTER || (T == RW_JOIN_LEFT_OUTER && S == 1) || (T == RW_JOIN_RIGHT_OUTER && S == 0); }
__device__ __host__ __forceinline__ bool jt_forward_exactly_once(int T, int S) { return ((T == RW_JOIN_LEFT_SEMI || T == RW_JOIN_LEFT_ANTI) && S == 0) || ((T == RW_JOIN_RIGHT_SEMI || T == RW_JOIN_RIGHT_ANTI) && S == 1); }
__device__ __host__ __forceinline__ bool jt_only_forward_matched_side(int T, int S) { return ((T == RW_JOIN_LEFT_SEMI || T == RW_JOIN_LEFT_ANTI) && S == 1) || ((T == RW_JOIN_RIGHT_SEMI || T == RW_JOIN_RIGHT_ANTI) && S == 0); }
__device__ __host__ __forceinline__ bool jt_is_semi(int T) { return T == RW_JOIN_LEFT_SEMI || T == RW_JOIN_RIGHT_SEMI; }
__device__ __host__ __forceinline__ bool jt_is_anti(int T) { return T == RW_JOIN_LEFT_ANTI || T == RW_JOIN_RIGHT_ANTI; }
__device__ __host__ __forceinline__ bool jt_forward_if_not_matched(int T, int S) { return (jt_is_anti(T) && jt_forward_exactly_once(T, S)) || jt_is_outer_side(T, S); }

// ------------------------------------------------------------------ key helpers
__device__ __forceinline__ bool chunk_key(const JoinPlanDev* p, int S, const DevChunk& ch, int64_t r, uint64_t* kw,
                                          uint32_t* nm) {
  uint32_t m = 0;
  bool never = false;
  for (int k = 0; k < p->n_keys; k++) {
    const ColRef& c = ch.cols[p->key_col[S][k]];
    if (col_is_null(c, r)) {
      m |= 1u << k;
      kw[k] = 0;
      if (!p->null_safe[k]) never = true;  // hash_join.rs:985-999
    } else {
      kw[k] = load_key_word(c, r);
    }
  }
  *nm = m;
  return never;
}

__device__ __forceinline__ uint64_t key_hash(const JoinPlanDev* p, const uint64_t* kw, uint32_t nm) {
  uint64_t h = 0x9e3779b97f4a7c15ull ^ nm;
  for (int k = 0; k < p->n_keys; k++) h = mix64(h ^ kw[k]) + 0x9e3779b97f4a7c15ull;
  return h;
}

// find the slot of a key (read-only). returns -1 if absent.
__device__ __forceinline__ int64_t js_find(const JoinSideDev& s, const JoinPlanDev* p, const uint64_t* kw, uint32_t nm) {
  const uint64_t mask = s.cap - 1;
  if (p->single_key) {
    if (nm) return (int64_t)s.cap;                 // NULL key side slot (null-safe equality)
    if (kw[0] == J_EMPTY) return (int64_t)s.cap + 1;
    uint64_t idx = mix64(kw[0]) & mask;
    while (true) {
      uint64_t cur = __ldcg((const unsigned long long*)(s.slots + idx * 2));
      if (cur == kw[0]) return (int64_t)idx;
      if (cur == J_EMPTY) return -1;
      idx = (idx + 1) & mask;
    }
  }
  uint64_t h = key_hash(p, kw, nm);
  uint64_t tag = (h & ~0xFFFFull) | ((uint64_t)nm << 8) | 1ull;
  uint64_t idx = (h >> 17) & mask;
  while (true) {
    const unsigned long long* ptr = (const unsigned long long*)(s.slots + idx * p->SW);
    unsigned long long cur = __ldcg(ptr);
    if (cur == 0ull) return -1;
    if ((cur & ~2ull) == tag) {
      while (cur & 2ull) cur = *(volatile const unsigned long long*)ptr;
      bool eq = true;
      for (int k = 0; k < p->n_keys; k++) eq = eq && (__ldcg(ptr + 1 + k) == kw[k]);
      if (eq) return (int64_t)idx;
    }
    idx = (idx + 1) & mask;
  }
}

__device__ __forceinline__ int64_t js_find_or_insert(const JoinSideDev& s, const JoinPlanDev* p, const uint64_t* kw,
                                                     uint32_t nm, bool* created) {
  const uint64_t mask = s.cap - 1;
  if (p->single_key) {
    if (nm) return (int64_t)s.cap;
    if (kw[0] == J_EMPTY) return (int64_t)s.cap + 1;
    uint64_t idx = mix64(kw[0]) & mask;
    while (true) {
      unsigned long long* ptr = (unsigned long long*)(s.slots + idx * 2);
      unsigned long long cur = __ldcg(ptr);
      if (cur == kw[0]) return (int64_t)idx;
      if (cur == J_EMPTY) {
        unsigned long long old = atomicCAS(ptr, (unsigned long long)J_EMPTY, (unsigned long long)kw[0]);
        if (old == J_EMPTY) { *created = true; return (int64_t)idx; }
        if (old == kw[0]) return (int64_t)idx;
      }
      idx = (idx + 1) & mask;
    }
  }
  uint64_t h = key_hash(p, kw, nm);
  uint64_t tag = (h & ~0xFFFFull) | ((uint64_t)nm << 8) | 1ull;
  uint64_t idx = (h >> 17) & mask;
  while (true) {
    unsigned long long* ptr = (unsigned long long*)(s.slots + idx * p->SW);
    unsigned long long cur = __ldcg(ptr);
    if (cur == 0ull) {
      unsigned long long old = atomicCAS(ptr, 0ull, (unsigned long long)(tag | 2ull));
      if (old == 0ull) {
        for (int k = 0; k < p->n_keys; k++) __stcg(ptr + 1 + k, (unsigned long long)kw[k]);
        __threadfence();
        atomicExch(ptr, (unsigned long long)tag);
        *created = true;
        return (int64_t)idx;
      }
      cur = old;
    }
    if ((cur & ~2ull) == tag) {
      while (cur & 2ull) cur = *(volatile unsigned long long*)ptr;
      bool eq = true;
      for (int k = 0; k < p->n_keys; k++) eq = eq && (__ldcg(ptr + 1 + k) == kw[k]);
      if (eq) return (int64_t)idx;
    }
    idx = (idx + 1) & mask;
  }
}

// head / count live in the last word of a slot: low 32 = head row (J_NIL = none), high 32 = live count
__device__ __forceinline__ uint32_t* slot_head(const JoinSideDev& s, const JoinPlanDev* p, int64_t slot) {
  return (uint32_t*)(s.slots + (uint64_t)slot * p->SW + p->KW);
}
__device__ __forceinline__ uint32_t* slot_count(const JoinSideDev& s, const JoinPlanDev* p, int64_t slot) {
  return slot_head(s, p, slot) + 1;
}

__global__ void join_init_slots_kernel(uint64_t* slots, uint64_t cap, int SW, int KW, int single_key) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap + 2; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t* s = slots + i * SW;
    s[0] = single_key ? J_EMPTY : 0ull;
    for (int k = 1; k < KW; k++) s[k] = 0;
    s[KW] = (uint64_t)J_NIL;  // head = NIL, count = 0
  }
}

// ------------------------------------------------------------------ row store access
__device__ __forceinline__ bool store_is_null(const JoinSideDev& s, int c, uint32_t row) {
  return s.valid[c] != nullptr && s.valid[c][row] == 0;
}
__device__ __forceinline__ uint64_t store_word(const JoinSideDev& s, const JoinPlanDev* p, int S, int c, uint32_t row) {
  ColRef cr;
  cr.data = s.col[c];
  cr.type = p->col_type[S][c];
  cr.width = p->col_width[S][c];
  return load_key_word(cr, row);
}

// copy one datum between columns of equal width
__device__ __forceinline__ void copy_datum(void* dst, int64_t di, const void* src, int64_t si, int width) {
  switch (width) {
    case 1: ((uint8_t*)dst)[di] = ((const uint8_t*)src)[si]; break;
    case 2: ((uint16_t*)dst)[di] = ((const uint16_t*)src)[si]; break;
    case 4: ((uint32_t*)dst)[di] = ((const uint32_t*)src)[si]; break;
    case 8: ((uint64_t*)dst)[di] = ((const uint64_t*)src)[si]; break;
    default: ((ulonglong2*)dst)[di] = ((const ulonglong2*)src)[si]; break;
  }
}

// JoinStreamChunkBuilder::{append_row, append_row_update, append_row_matched}  builder.rs:84-148
__device__ __forceinline__ void emit_row(const JoinOutDev& o, const JoinPlanDev* p, int64_t orow, uint8_t op, int S,
                                         const DevChunk& ch, int64_t ur, const JoinSideDev& ms, int64_t mr) {
  o.ops[orow] = op;
  o.vis[orow] = 1;
  const int n_u = p->n_map[S], n_m = p->n_map[1 - S];
  for (int i = 0; i < n_u; i++) {
    const int ic = p->map_in[S][i], oc = p->map_out[S][i];
    bool nul = true;
    if (ur >= 0) {
      const ColRef& c = ch.cols[ic];
      nul = col_is_null(c, ur);
      if (!nul) copy_datum(o.col[oc], orow, c.data, ur, c.width);
    }
    o.valid[oc][orow] = nul ? 0 : 1;
    if (nul) o.has_null[oc] = 1;
  }
  for (int i = 0; i < n_m; i++) {
    const int ic = p->map_in[1 - S][i], oc = p->map_out[1 - S][i];
    bool nul = true;
    if (mr >= 0) {
      nul = store_is_null(ms, ic, (uint32_t)mr);
      if (!nul) copy_datum(o.col[oc], orow, ms.col[ic], mr, p->col_width[1 - S][ic]);
    }
    o.valid[oc][orow] = nul ? 0 : 1;
    if (nul) o.has_null[oc] = 1;
  }
}

// check_join_condition (hash_join.rs:1362-1384) restricted to one integer comparison
__device__ __forceinline__ bool cond_ok(const JoinPlanDev* p, int S, const DevChunk& ch, int64_t ur, const JoinSideDev& ms,
                                        uint32_t mr) {
  if (p->cond_cmp == RW_CMP_NONE) return true;
  const int nl = p->n_cols[0];
  int64_t v[2];
  const int idx[2] = {p->cond_lhs, p->cond_rhs};
  for (int t = 0; t < 2; t++) {
    const bool left = idx[t] < nl;
    const int local = left ? idx[t] : idx[t] - nl;
    const int side = left ? 0 : 1;
    if (side == S) {
      if (col_is_null(ch.cols[local], ur)) return false;
      v[t] = load_i64(ch.cols[local], ur);
    } else {
      if (store_is_null(ms, local, mr)) return false;
      v[t] = (int64_t)store_word(ms, p, 1 - S, local, mr);
    }
  }
  switch (p->cond_cmp) {
    case RW_CMP_LT: return v[0] < v[1];
    case RW_CMP_LE: return v[0] <= v[1];
    case RW_CMP_GT: return v[0] > v[1];
    case RW_CMP_GE: return v[0] >= v[1];
    case RW_CMP_EQ: return v[0] == v[1];
    default: return v[0] != v[1];
  }
}

// does stored row `row` of side S carry the same pk as chunk row r ?  (pk = deduped_pk_indices;
// the join key is equal by construction: join/hash_join.rs:710-713)
__device__ __forceinline__ bool pk_equal(const JoinPlanDev* p, int S, const JoinSideDev& s, uint32_t row, const DevChunk& ch,
                                         int64_t r) {
  for (int i = 0; i < p->n_pk[S]; i++) {
    const int c = p->pk_col[S][i];
    const bool n1 = store_is_null(s, c, row), n2 = col_is_null(ch.cols[c], r);
    if (n1 != n2) return false;
    if (n1) continue;
    if (p->col_width[S][c] == 16) {
      const uint64_t* a = (const uint64_t*)s.col[c] + (uint64_t)row * 2;
      const uint64_t* b = (const uint64_t*)ch.cols[c].data + r * 2;
      if (a[0] != b[0] || a[1] != b[1]) return false;
    } else if (store_word(s, p, S, c, row) != load_key_word(ch.cols[c], r)) {
      return false;
    }
  }
  return true;
}

// write chunk row r into the store at `row`
__device__ __forceinline__ void store_write_row(const JoinPlanDev* p, int S, const JoinSideDev& s, uint32_t row,
                                                const DevChunk& ch, int64_t r) {
  for (int c = 0; c < p->n_cols[S]; c++) {
    const ColRef& cr = ch.cols[c];
    const bool nul = col_is_null(cr, r);
    if (s.valid[c]) s.valid[c][row] = nul ? 0 : 1;
    if (!nul) copy_datum(s.col[c], row, cr.data, r, cr.width);
  }
}

// =============================================================================== generic path
struct JoinScratch {
  uint64_t* sortkey;      // [n]  gid << 32 | row
  uint64_t* sortkey_alt;  // [n]
  uint64_t* packed;       // [n]  store_flag << 40 | out bound
  uint64_t* offs;         // [n]  exclusive scan of packed
  int64_t* match_slot;    // [n]
  int32_t* gtable;        // [gcap] batch-local key -> representative row
  uint64_t gcap;
};

// G1: per row -- never-match rule, probe of the other side's index, output bound, batch-local group id
__global__ void __launch_bounds__(256) join_prepare_kernel(const JoinPlanDev* __restrict__ p, int S, DevChunk ch,
                                                            JoinSideDev other, JoinScratch sc) {
  const int T = p->T;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < ch.n; r += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t op = ch.ops[r];
    if (!row_visible(ch, r, op)) {
      sc.sortkey[r] = 0xFFFFFFFF00000000ull | (uint64_t)r;
      sc.packed[r] = 0;
      sc.match_slot[r] = -1;
      continue;
    }
    uint64_t kw[RW_MAX_KEYS];
    uint32_t nm;
    const bool never = chunk_key(p, S, ch, r, kw, &nm);
    const bool ins = (op == RW_OP_INSERT || op == RW_OP_UPDATE_INSERT);
    uint32_t gid;
    int64_t ms = -1;
    uint64_t bound;
    if (never) {
      gid = 0x80000000u | (uint32_t)r;  // singleton group
      bound = jt_forward_if_not_matched(T, S) ? 1 : 0;
    } else {
      ms = js_find(other, p, kw, nm);
      const uint64_t m = ms >= 0 ? (uint64_t)*slot_count(other, p, ms) : 0;
      uint64_t per_match;
      if (T == RW_JOIN_INNER) per_match = 1;
      else if (jt_is_semi(T) || jt_is_anti(T)) per_match = jt_forward_exactly_once(T, S) ? 0 : 1;
      else per_match = jt_outer_side_null(T, S) ? 2 : 1;
      const uint64_t fwd = (jt_forward_if_not_matched(T, S) || (jt_is_semi(T) && jt_forward_exactly_once(T, S))) ? 1 : 0;
      bound = per_match * m + fwd;
      // batch-local grouping: claim a scratch slot with this row as representative, or join the
      // group whose representative carries an equal key
      const uint64_t gmask = sc.gcap - 1;
      uint64_t gi = key_hash(p, kw, nm) & gmask;
      while (true) {
        int cur = sc.gtable[gi];
        if (cur < 0) {
          int old = atomicCAS(sc.gtable + gi, -1, (int)r);
          if (old < 0) break;
          cur = old;
        }
        uint64_t kw2[RW_MAX_KEYS];
        uint32_t nm2;
        chunk_key(p, S, ch, cur, kw2, &nm2);
        bool eq = (nm2 == nm);
        for (int k = 0; k < p->n_keys; k++) eq = eq && (kw2[k] == kw[k]);
        if (eq) break;
        gi = (gi + 1) & gmask;
      }
      gid = (uint32_t)gi;
    }
    sc.sortkey[r] = ((uint64_t)gid << 32) | (uint64_t)r;
    sc.packed[r] = ((uint64_t)((ins && !never) ? 1 : 0) << 40) | bound;
    sc.match_slot[r] = ms;
  }
}

__global__ void join_totals_kernel(const uint64_t* packed, const uint64_t* offs, int64_t n, JoinStatus* st) {
  if (n > 0) {
    uint64_t tot = offs[n - 1] + packed[n - 1];
    st->out_rows = tot & ((1ull << 40) - 1);
    st->n_store = tot >> 40;
  } else {
    st->out_rows = 0;
    st->n_store = 0;
  }
}

__global__ void fill_i32_kernel(int32_t* p, uint64_t n, int32_t v) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// G4: one thread per join key of the batch, rows in input order (exact sequential semantics per key)
__global__ void __launch_bounds__(128) join_serial_kernel(const JoinPlanDev* __restrict__ p, int S, DevChunk ch,
                                                           JoinSideDev own, JoinSideDev other, JoinScratch sc,
                                                           const uint64_t* __restrict__ sorted, JoinOutDev o,
                                                           JoinStatus* st, uint32_t store_base) {
  const int T = p->T;
  const bool fwd_once = jt_forward_exactly_once(T, S);
  const bool fwd_unmatched = jt_forward_if_not_matched(T, S);
  const bool fwd_matched = jt_is_semi(T) && fwd_once;
  const bool only_matched = jt_only_forward_matched_side(T, S);
  const bool side_null = jt_outer_side_null(T, S);
  const bool other_deg = other.degree != nullptr;
  unsigned int new_keys = 0;
  long long live_own = 0, live_other = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ch.n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t sk = sorted[i];
    const uint32_t gid = (uint32_t)(sk >> 32);
    if (gid == 0xFFFFFFFFu) continue;                          // invisible rows
    if (i > 0 && (uint32_t)(sorted[i - 1] >> 32) == gid) continue;  // not a group start
    int64_t own_slot = -2;  // lazily resolved
    for (int64_t j = i; j < ch.n && (uint32_t)(sorted[j] >> 32) == gid; j++) {
      const int64_t r = (int64_t)(sorted[j] & 0xFFFFFFFFull);
      const uint8_t op = ch.ops[r];
      const bool ins = (op == RW_OP_INSERT || op == RW_OP_UPDATE_INSERT);
      const uint8_t jop = ins ? RW_OP_INSERT : RW_OP_DELETE;
      const uint64_t pk = sc.packed[r];
      const int64_t bound = (int64_t)(pk & ((1ull << 40) - 1));
      const int64_t obase = (int64_t)(sc.offs[r] & ((1ull << 40) - 1));
      const uint32_t store_row = store_base + (uint32_t)(sc.offs[r] >> 40);
      int64_t w = 0;  // rows written so far for r
      const bool room = obase + bound <= o.capacity;
      if (!room) atomicOr(&st->err, JERR_OUT_CAPACITY);
      if (gid & 0x80000000u) {  // CacheResult::NeverMatch (hash_join.rs:1126-1135): forwarded, never stored
        if (fwd_unmatched && room) emit_row(o, p, obase + w++, jop, S, ch, r, other, -1);
        continue;
      }
      const int64_t ms = sc.match_slot[r];
      uint32_t degree = 0;
      int64_t ao_row = -1;
      if (ms >= 0) {
        uint32_t m = *slot_head(other, p, ms) & 0x7fffffffu;
        while (m != J_NIL) {
          const uint32_t lk = other.link[m];
          if (!(lk & J_DEAD)) {
            if (cond_ok(p, S, ch, r, other, m)) {
              degree++;
              uint32_t md = other_deg ? other.degree[m] : 0;
              if (ins && !fwd_once && room) {  // with_match_on_insert (builder.rs:184-231): m.degree BEFORE the increment
                if (jt_is_anti(T)) { if (md == 0 && only_matched) emit_row(o, p, obase + w++, RW_OP_DELETE, S, ch, -1, other, m); }
                else if (jt_is_semi(T)) { if (md == 0 && only_matched) emit_row(o, p, obase + w++, RW_OP_INSERT, S, ch, -1, other, m); }
                else if (md == 0 && side_null) {
                  emit_row(o, p, obase + w++, RW_OP_DELETE, S, ch, -1, other, m);
                  emit_row(o, p, obase + w++, RW_OP_INSERT, S, ch, r, other, m);
                } else emit_row(o, p, obase + w++, RW_OP_INSERT, S, ch, r, other, m);
              }
              if (other_deg) { md = ins ? md + 1 : md - 1; other.degree[m] = md; }  // update_degree (join/hash_join.rs:355-380)
              if (!ins && !fwd_once && room) {  // with_match_on_delete (builder.rs:233-284): m.degree AFTER the decrement
                if (jt_is_anti(T)) { if (md == 0 && only_matched) emit_row(o, p, obase + w++, RW_OP_INSERT, S, ch, -1, other, m); }
                else if (jt_is_semi(T)) { if (md == 0 && only_matched) emit_row(o, p, obase + w++, RW_OP_DELETE, S, ch, -1, other, m); }
                else if (md == 0 && side_null) {
                  emit_row(o, p, obase + w++, RW_OP_DELETE, S, ch, r, other, m);
                  emit_row(o, p, obase + w++, RW_OP_INSERT, S, ch, -1, other, m);
                } else emit_row(o, p, obase + w++, RW_OP_DELETE, S, ch, r, other, m);
              }
            }
            if (p->append_only_optimize) {  // hash_join.rs:1339-1345 (regardless of the condition)
              if (ao_row >= 0) atomicOr(&st->err, JERR_APPEND_ONLY_MULTI);
              ao_row = m;
            }
          }
          m = lk & 0x7fffffffu;
        }
      }
      // forward rows depending on join types (hash_join.rs:1198-1210)
      if (room) {
        if (degree == 0) { if (fwd_unmatched) emit_row(o, p, obase + w++, jop, S, ch, r, other, -1); }
        else if (fwd_matched) emit_row(o, p, obase + w++, jop, S, ch, r, other, -1);
        for (; w < bound; w++) {  // unused reserved rows become invisible holes
          o.ops[obase + w] = RW_OP_INSERT;
          o.vis[obase + w] = 0;
          o.has_null[J_MAX_OUT] = 1;
          for (int k = 0; k < p->n_out; k++) o.valid[k][obase + w] = 0;
        }
      }
      // append-only optimisation (hash_join.rs:1222-1228): drop the matched row, do not store u
      if (p->append_only_optimize && ao_row >= 0) {
        other.link[ao_row] |= J_DEAD;
        *slot_count(other, p, ms) -= 1;
        live_other--;
        continue;
      }
      // own-side state (hash_join.rs:1230-1242; JoinHashMap::insert / delete join/hash_join.rs:591-681)
      if (own_slot == -2) {
        uint64_t kw[RW_MAX_KEYS];
        uint32_t nm;
        chunk_key(p, S, ch, r, kw, &nm);
        bool created = false;
        own_slot = ins ? js_find_or_insert(own, p, kw, nm, &created) : js_find(own, p, kw, nm);
        if (created) new_keys++;
        if (own_slot < 0 && ins) own_slot = -2;
      }
      if (ins) {
        store_write_row(p, S, own, store_row, ch, r);
        if (own.degree) own.degree[store_row] = degree;
        uint32_t* hd = slot_head(own, p, own_slot);
        own.link[store_row] = *hd & 0x7fffffffu;
        *hd = store_row;
        *slot_count(own, p, own_slot) += 1;
        live_own++;
      } else {
        bool found = false;
        if (own_slot >= 0) {
          uint32_t m = *slot_head(own, p, own_slot) & 0x7fffffffu;
          while (m != J_NIL) {
            const uint32_t lk = own.link[m];
            if (!(lk & J_DEAD) && pk_equal(p, S, own, m, ch, r)) {
              own.link[m] = lk | J_DEAD;
              *slot_count(own, p, own_slot) -= 1;
              live_own--;
              found = true;
              break;
            }
            m = lk & 0x7fffffffu;
          }
        } else {
          own_slot = -2;  // key may be created by a later insert of this group
        }
        if (!found && p->strict) atomicOr(&st->err, JERR_DOUBLE_DELETE);
      }
    }
  }
  if (new_keys) atomicAdd(&st->n_keys[S], (unsigned long long)new_keys);
  if (live_own) atomicAdd(&st->live_rows[S], (unsigned long long)live_own);
  if (live_other) atomicAdd(&st->live_rows[1 - S], (unsigned long long)live_other);
}

// =============================================================================== inner fast path
// F1: probe + emit, fused.  Each thread walks the matched chain once, buffering up to 4 matches in
// registers; blocks reserve output ranges with one atomicAdd after a block-wide scan (warp shuffles).
// The kernel does not mutate operator state, so it is simply re-run with a larger output buffer
// if the reservation overflowed.
#define JF_BLOCK 256
__global__ void __launch_bounds__(JF_BLOCK) join_inner_probe_emit_kernel(const JoinPlanDev* __restrict__ p, int S,
                                                                          DevChunk ch, JoinSideDev other, JoinOutDev o,
                                                                          JoinStatus* st, uint32_t* store_flag /* [n] */) {
  __shared__ unsigned long long warp_tot[JF_BLOCK / 32];
  __shared__ unsigned long long block_base;
  const int lane = lane_id(), wid = threadIdx.x >> 5;
  const int64_t n_iter = (ch.n + (int64_t)gridDim.x * JF_BLOCK - 1) / ((int64_t)gridDim.x * JF_BLOCK);
  for (int64_t it = 0; it < n_iter; it++) {
    const int64_t r = (it * gridDim.x + blockIdx.x) * (int64_t)JF_BLOCK + threadIdx.x;
    uint32_t cnt = 0, first[4], head = J_NIL;
    uint8_t op = 0;
    if (r < ch.n) {
      op = ch.ops[r];
      uint32_t sf = 0;
      if (row_visible(ch, r, op)) {
        uint64_t kw[RW_MAX_KEYS];
        uint32_t nm;
        if (!chunk_key(p, S, ch, r, kw, &nm)) {
          sf = (op == RW_OP_INSERT || op == RW_OP_UPDATE_INSERT) ? 1u : 0u;
          const int64_t ms = js_find(other, p, kw, nm);
          if (ms >= 0) {
            head = *slot_head(other, p, ms) & 0x7fffffffu;
            uint32_t m = head;
            while (m != J_NIL) {
              const uint32_t lk = __ldcg(other.link + m);
              if (!(lk & J_DEAD) && cond_ok(p, S, ch, r, other, m)) {
                if (cnt < 4) first[cnt] = m;
                cnt++;
              }
              m = lk & 0x7fffffffu;
            }
          }
        }
      }
      store_flag[r] = sf;
    }
    // block-wide exclusive scan of cnt
    unsigned long long incl = cnt;
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long run = 0;
      for (int w = 0; w < JF_BLOCK / 32; w++) { unsigned long long t = warp_tot[w]; warp_tot[w] = run; run += t; }
      block_base = run ? atomicAdd(&st->out_rows, run) : 0ull;
    }
    __syncthreads();
    if (cnt) {
      int64_t pos = (int64_t)(block_base + warp_tot[wid] + incl - cnt);
      if (pos + cnt > o.capacity) {
        atomicOr(&st->err, JERR_OUT_CAPACITY);
      } else {
        const uint8_t oop = (op == RW_OP_INSERT || op == RW_OP_UPDATE_INSERT) ? RW_OP_INSERT : RW_OP_DELETE;
        if (cnt <= 4) {
          for (uint32_t k = 0; k < cnt; k++) emit_row(o, p, pos + k, oop, S, ch, r, other, first[k]);
        } else {
          uint32_t m = head;
          while (m != J_NIL) {
            const uint32_t lk = __ldcg(other.link + m);
            if (!(lk & J_DEAD) && cond_ok(p, S, ch, r, other, m)) emit_row(o, p, pos++, oop, S, ch, r, other, m);
            m = lk & 0x7fffffffu;
          }
        }
      }
    }
    __syncthreads();
  }
}

// F2: own-side inserts (row-parallel).  Store row ids are store_base + (number of stored rows before r)
// -- an exclusive scan of the store flags -- so ids increase with the chunk position.
__global__ void __launch_bounds__(256) join_inner_insert_kernel(const JoinPlanDev* __restrict__ p, int S, DevChunk ch,
                                                                 JoinSideDev own, JoinStatus* st, uint32_t store_base,
                                                                 const uint32_t* __restrict__ store_flag,
                                                                 const uint32_t* __restrict__ store_rank) {
  unsigned int new_keys = 0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < ch.n; r += (int64_t)gridDim.x * blockDim.x) {
    if (r == ch.n - 1) st->n_store = (unsigned long long)store_rank[r] + store_flag[r];
    if (!store_flag[r]) continue;
    uint64_t kw[RW_MAX_KEYS];
    uint32_t nm = 0;
    chunk_key(p, S, ch, r, kw, &nm);
    const uint32_t row = store_base + store_rank[r];
    store_write_row(p, S, own, row, ch, r);
    bool created = false;
    const int64_t slot = js_find_or_insert(own, p, kw, nm, &created);
    if (created) new_keys++;
    const uint32_t old = atomicExch(slot_head(own, p, slot), row);
    own.link[row] = old & 0x7fffffffu;
    atomicAdd(slot_count(own, p, slot), 1u);
  }
  for (int d = 16; d > 0; d >>= 1) new_keys += __shfl_xor_sync(0xffffffffu, new_keys, d);
  if (lane_id() == 0 && new_keys) atomicAdd(&st->n_keys[S], (unsigned long long)new_keys);
}

// F3: own-side deletes (row-parallel, after F2).  Sequential rule: the delete at chunk position r
// removes the live row with equal pk that was inserted most recently BEFORE position r, i.e. the
// largest store row id below store_base + rank(r).
__global__ void __launch_bounds__(256) join_inner_delete_kernel(const JoinPlanDev* __restrict__ p, int S, DevChunk ch,
                                                                 JoinSideDev own, JoinStatus* st, uint32_t store_base,
                                                                 const uint32_t* __restrict__ store_rank /* [n] */) {
  long long removed = 0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < ch.n; r += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t op = ch.ops[r];
    if (!row_visible(ch, r, op) || !(op == RW_OP_DELETE || op == RW_OP_UPDATE_DELETE)) continue;
    uint64_t kw[RW_MAX_KEYS];
    uint32_t nm;
    if (chunk_key(p, S, ch, r, kw, &nm)) continue;  // never-match rows were never stored
    const int64_t slot = js_find(own, p, kw, nm);
    bool found = false;
    if (slot >= 0) {
      const uint32_t bound = store_base + store_rank[r];
      while (!found) {
        uint32_t best = J_NIL;
        uint32_t m = *slot_head(own, p, slot) & 0x7fffffffu;
        while (m != J_NIL) {
          const uint32_t lk = __ldcg(own.link + m);
          if (!(lk & J_DEAD) && m < bound && (best == J_NIL || m > best) && pk_equal(p, S, own, m, ch, r)) best = m;
          m = lk & 0x7fffffffu;
        }
        if (best == J_NIL) break;
        const uint32_t old = atomicOr(own.link + best, J_DEAD);
        if (!(old & J_DEAD)) {
          atomicSub(slot_count(own, p, slot), 1u);
          removed++;
          found = true;
        }
      }
    }
    if (!found && p->strict) atomicOr(&st->err, JERR_DOUBLE_DELETE);
  }
  if (removed) atomicAdd(&st->live_rows[S], (unsigned long long)(-removed));
}

// ------------------------------------------------------------------ growth helpers
__global__ void join_rehash_kernel(const uint64_t* os, uint64_t ocap, uint64_t* ns, uint64_t ncap, int SW, int KW,
                                   int single_key, int n_keys) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < ocap + 2; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t* s = os + i * SW;
    uint64_t dst;
    if (i >= ocap) {
      dst = ncap + (i - ocap);
    } else {
      const uint64_t w0 = s[0];
      if (single_key ? (w0 == J_EMPTY) : (w0 == 0)) continue;
      const uint64_t mask = ncap - 1;
      uint64_t idx;
      if (single_key) {
        idx = mix64(w0) & mask;
        while (atomicCAS((unsigned long long*)(ns + idx * SW), (unsigned long long)J_EMPTY, (unsigned long long)w0) != J_EMPTY) idx = (idx + 1) & mask;
      } else {
        uint32_t nm = (uint32_t)((w0 >> 8) & 0xff);
        uint64_t h = 0x9e3779b97f4a7c15ull ^ nm;
        for (int k = 0; k < n_keys; k++) h = mix64(h ^ s[1 + k]) + 0x9e3779b97f4a7c15ull;
        idx = (h >> 17) & mask;
        while (atomicCAS((unsigned long long*)(ns + idx * SW), 0ull, (unsigned long long)w0) != 0ull) idx = (idx + 1) & mask;
      }
      dst = idx;
    }
    for (int k = (i >= ocap ? 0 : 1); k < SW; k++) ns[dst * SW + k] = s[k];
  }
}

static inline size_t align_up_j(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace rw

// =============================================================================== host handle
using namespace rw;

struct JoinSideHost {
  int n_cols = 0;
  std::vector<int> types;
  DevBuf col[RW_MAX_COLS], valid[RW_MAX_COLS], link, degree, slots;
  bool has_valid[RW_MAX_COLS];
  bool need_degree = false;
  uint64_t row_cap = 0;   // rows allocated in the store
  uint64_t n_rows = 0;    // rows handed out (incl. dead / cancelled)
  uint64_t slot_cap = 0;
  uint64_t keys_upper = 0;
};

struct rwgpu_join {
  JoinPlanDev plan;
  DevBuf plan_dev, status;
  PinnedBuf status_host;
  JoinSideHost side[2];
  cudaStream_t stream = nullptr;
  std::vector<int> out_types;
  int chunk_size = 1024;
  bool fast_inner = false;
  uint64_t launches = 0;
  KernelProf prof;
  // scratch
  DevBuf sk, sk_alt, packed, offs, mslot, gtable, cub_tmp, row_of, row_rev, row_bound;
  int64_t scratch_rows = 0;
  uint64_t gcap = 0;
  size_t cub_bytes = 0;
  // output (device)
  DevBuf out_ops, out_vis, out_hasnull, out_col[J_MAX_OUT], out_valid[J_MAX_OUT], out_bits[J_MAX_OUT], out_visbits;
  int64_t out_cap = 0;
  // host upload staging
  DevBuf up;
  PinnedBuf up_host;
  std::vector<rw_column> dev_view_cols;
  ~rwgpu_join() { if (stream) cudaStreamDestroy(stream); }
};

static int jgrid(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, 148 * 8));
}

static JoinSideDev side_dev(const rwgpu_join* h, int S) {
  const JoinSideHost& s = h->side[S];
  JoinSideDev d;
  memset(&d, 0, sizeof(d));
  for (int c = 0; c < s.n_cols; c++) {
    d.col[c] = s.col[c].p;
    d.valid[c] = s.has_valid[c] ? s.valid[c].as<uint8_t>() : nullptr;
  }
  d.link = s.link.as<uint32_t>();
  d.degree = s.need_degree ? s.degree.as<uint32_t>() : nullptr;
  d.slots = s.slots.as<uint64_t>();
  d.cap = s.slot_cap;
  return d;
}

static int join_alloc_slots(rwgpu_join* h, DevBuf& buf, uint64_t cap) {
  RW_CUDA(buf.reserve((cap + 2) * h->plan.SW * 8));
  join_init_slots_kernel<<<jgrid((int64_t)cap + 2, 256), 256, 0, h->stream>>>(buf.as<uint64_t>(), cap, h->plan.SW, h->plan.KW, h->plan.single_key);
  RW_CUDA(cudaGetLastError());
  h->launches++;
  return RW_OK;
}

// grow the row store of side S to hold at least `rows` rows (contents preserved)
static int join_grow_store(rwgpu_join* h, int S, uint64_t rows) {
  JoinSideHost& s = h->side[S];
  if (rows <= s.row_cap) return RW_OK;
  if (rows >= 0x7ffffff0ull) return fail(RW_ERR_OOM, "join side exceeds 2^31 rows");
  uint64_t ncap = std::max<uint64_t>(s.row_cap * 2, std::max<uint64_t>(rows, 1 << 16));
  ncap = std::min<uint64_t>(ncap, 0x7ffffff0ull);
  auto grow = [&](DevBuf& b, size_t elt) -> int {
    DevBuf nb;
    RW_CUDA(nb.reserve(ncap * elt));
    if (s.n_rows) RW_CUDA(cudaMemcpyAsync(nb.p, b.p, s.n_rows * elt, cudaMemcpyDeviceToDevice, h->stream));
    RW_CUDA(cudaStreamSynchronize(h->stream));
    b = std::move(nb);
    return RW_OK;
  };
  for (int c = 0; c < s.n_cols; c++) {
    int rc = grow(s.col[c], (size_t)type_width(s.types[c]));
    if (rc != RW_OK) return rc;
    if (s.has_valid[c]) { rc = grow(s.valid[c], 1); if (rc != RW_OK) return rc; }
  }
  int rc = grow(s.link, 4);
  if (rc != RW_OK) return rc;
  if (s.need_degree) { rc = grow(s.degree, 4); if (rc != RW_OK) return rc; }
  s.row_cap = ncap;
  return RW_OK;
}

// a column of side S is about to receive NULLs for the first time: materialise its valid bytes
static int join_enable_valid(rwgpu_join* h, int S, int c) {
  JoinSideHost& s = h->side[S];
  if (s.has_valid[c]) return RW_OK;
  RW_CUDA(s.valid[c].reserve(std::max<uint64_t>(s.row_cap, 1)));
  RW_CUDA(cudaMemsetAsync(s.valid[c].p, 1, std::max<uint64_t>(s.row_cap, 1), h->stream));
  s.has_valid[c] = true;
  return RW_OK;
}

static int join_grow_slots(rwgpu_join* h, int S, uint64_t need_keys) {
  JoinSideHost& s = h->side[S];
  if (need_keys * 2 <= s.slot_cap) return RW_OK;
  uint64_t ncap = s.slot_cap;
  while (ncap < need_keys * 4) ncap <<= 1;
  DevBuf nb;
  int rc = join_alloc_slots(h, nb, ncap);
  if (rc != RW_OK) return rc;
  join_rehash_kernel<<<jgrid((int64_t)s.slot_cap + 2, 256), 256, 0, h->stream>>>(s.slots.as<uint64_t>(), s.slot_cap, nb.as<uint64_t>(), ncap,
                                                                                  h->plan.SW, h->plan.KW, h->plan.single_key, h->plan.n_keys);
  RW_CUDA(cudaGetLastError());
  h->launches++;
  RW_CUDA(cudaStreamSynchronize(h->stream));
  s.slots = std::move(nb);
  s.slot_cap = ncap;
  return RW_OK;
}

static int join_ensure_scratch(rwgpu_join* h, int64_t n) {
  if (n > h->scratch_rows) {
    int64_t cap = std::max<int64_t>(n, 4096);
    RW_CUDA(h->sk.reserve((size_t)cap * 8));
    RW_CUDA(h->sk_alt.reserve((size_t)cap * 8));
    RW_CUDA(h->packed.reserve((size_t)cap * 8));
    RW_CUDA(h->offs.reserve((size_t)cap * 8));
    RW_CUDA(h->mslot.reserve((size_t)cap * 8));
    RW_CUDA(h->row_of.reserve((size_t)cap * 4));
    RW_CUDA(h->row_rev.reserve((size_t)cap * 4));
    RW_CUDA(h->row_bound.reserve((size_t)cap * 4));
    uint64_t g = 1024;
    while (g < (uint64_t)cap * 2) g <<= 1;
    RW_CUDA(h->gtable.reserve(g * 4));
    h->gcap = g;
    size_t b1 = 0, b2 = 0, b3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b1, (uint64_t*)nullptr, (uint64_t*)nullptr, (int)cap);
    cub::DoubleBuffer<uint64_t> db((uint64_t*)nullptr, (uint64_t*)nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, b2, db, (int)cap);
    cub::DeviceScan::ExclusiveSum(nullptr, b3, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)cap);
    h->cub_bytes = std::max(b1, std::max(b2, b3)) + 256;
    RW_CUDA(h->cub_tmp.reserve(h->cub_bytes));
    h->scratch_rows = cap;
  }
  return RW_OK;
}

static int join_ensure_out(rwgpu_join* h, int64_t rows) {
  if (rows <= h->out_cap) return RW_OK;
  int64_t cap = std::max<int64_t>(rows + rows / 4, 4096);
  RW_CUDA(h->out_ops.reserve((size_t)cap));
  RW_CUDA(h->out_vis.reserve((size_t)cap));
  RW_CUDA(h->out_visbits.reserve((size_t)((cap + 63) / 64) * 8));
  for (size_t k = 0; k < h->out_types.size(); k++) {
    RW_CUDA(h->out_col[k].reserve((size_t)cap * type_width(h->out_types[k])));
    RW_CUDA(h->out_valid[k].reserve((size_t)cap));
    RW_CUDA(h->out_bits[k].reserve((size_t)((cap + 63) / 64) * 8));
  }
  h->out_cap = cap;
  return RW_OK;
}

static JoinOutDev out_dev(rwgpu_join* h) {
  JoinOutDev o;
  memset(&o, 0, sizeof(o));
  o.ops = h->out_ops.as<uint8_t>();
  o.vis = h->out_vis.as<uint8_t>();
  for (size_t k = 0; k < h->out_types.size(); k++) { o.col[k] = h->out_col[k].p; o.valid[k] = h->out_valid[k].as<uint8_t>(); }
  o.has_null = h->out_hasnull.as<unsigned int>();
  o.capacity = h->out_cap;
  return o;
}

static int join_read_status(rwgpu_join* h, cudaStream_t st, JoinStatus* out) {
  RW_CUDA(cudaMemcpyAsync(h->status_host.p, h->status.p, sizeof(JoinStatus), cudaMemcpyDeviceToHost, st));
  RW_CUDA(cudaStreamSynchronize(st));
  memcpy(out, h->status_host.p, sizeof(JoinStatus));
  return RW_OK;
}

static int join_check_err(rwgpu_join* h, const JoinStatus& s, cudaStream_t st) {
  if (!s.err) return RW_OK;
  unsigned int e = s.err;
  cudaMemsetAsync(&h->status.as<JoinStatus>()->err, 0, sizeof(unsigned int), st);
  if (e & JERR_DOUBLE_DELETE) return fail(RW_ERR_INCONSISTENT, "removing a join state entry but it is not in the cache");
  if (e & JERR_APPEND_ONLY_MULTI) return fail(RW_ERR_INCONSISTENT, "append-only optimisation: more than one matched row");
  return fail(RW_ERR_CUDA, "internal: join output capacity");
}

// one push of a device-resident chunk; on return the output sits in the device output buffers
static int join_push_dev(rwgpu_join* h, int S, const DevChunk& ch, cudaStream_t st, int64_t* out_rows,
                         unsigned int* has_null_host) {
  *out_rows = 0;
  memset(has_null_host, 0, sizeof(unsigned int) * (J_MAX_OUT + 1));
  const int64_t n = ch.n;
  if (n <= 0) return RW_OK;
  if (n >= (1ll << 31)) return fail(RW_ERR_INVALID, "chunk too large");
  JoinSideHost& own = h->side[S];
  // NULL-carrying input columns need valid bytes in the store
  for (int c = 0; c < own.n_cols; c++)
    if (ch.cols[c].valid_bits || ch.cols[c].valid_bytes) { int rc = join_enable_valid(h, S, c); if (rc != RW_OK) return rc; }
  int rc = join_ensure_scratch(h, n);
  if (rc != RW_OK) return rc;
  rc = join_grow_store(h, S, own.n_rows + (uint64_t)n);
  if (rc != RW_OK) return rc;
  rc = join_grow_slots(h, S, own.keys_upper + (uint64_t)n);
  if (rc != RW_OK) return rc;
  JoinStatus* ds = h->status.as<JoinStatus>();
  const JoinPlanDev* pd = h->plan_dev.as<JoinPlanDev>();
  RW_CUDA(cudaMemsetAsync(ds, 0, 16, st));  // out_rows, n_store
  RW_CUDA(cudaMemsetAsync(h->out_hasnull.p, 0, sizeof(unsigned int) * (J_MAX_OUT + 1), st));
  JoinStatus hs;
  if (h->fast_inner) {
    rc = join_ensure_out(h, std::max<int64_t>(2 * n, 4096));
    if (rc != RW_OK) return rc;
    while (true) {
      h->prof.begin(st);
      join_inner_probe_emit_kernel<<<jgrid(n, JF_BLOCK), JF_BLOCK, 0, st>>>(pd, S, ch, side_dev(h, 1 - S), out_dev(h), ds,
                                                                             h->row_of.as<uint32_t>());
      h->prof.end(st);
      RW_CUDA(cudaGetLastError());
      h->launches++;
      rc = join_read_status(h, st, &hs);
      if (rc != RW_OK) return rc;
      if (!(hs.err & JERR_OUT_CAPACITY)) break;
      // overflow: the probe kernel is read-only, re-run it with room for every reserved row
      RW_CUDA(cudaMemsetAsync(ds, 0, 16, st));
      RW_CUDA(cudaMemsetAsync(&ds->err, 0, 4, st));
      RW_CUDA(cudaMemsetAsync(h->out_hasnull.p, 0, sizeof(unsigned int) * (J_MAX_OUT + 1), st));
      rc = join_ensure_out(h, (int64_t)hs.out_rows);
      if (rc != RW_OK) return rc;
    }
    const int64_t produced = (int64_t)hs.out_rows;
    JoinSideDev od = side_dev(h, S);
    size_t tb = h->cub_bytes;
    cub::DeviceScan::ExclusiveSum(h->cub_tmp.p, tb, h->row_of.as<uint32_t>(), h->row_rev.as<uint32_t>(), (int)n, st);
    join_inner_insert_kernel<<<jgrid(n, 256), 256, 0, st>>>(pd, S, ch, od, ds, (uint32_t)own.n_rows, h->row_of.as<uint32_t>(),
                                                            h->row_rev.as<uint32_t>());
    join_inner_delete_kernel<<<jgrid(n, 256), 256, 0, st>>>(pd, S, ch, od, ds, (uint32_t)own.n_rows, h->row_rev.as<uint32_t>());
    RW_CUDA(cudaGetLastError());
    h->launches += 3;
    rc = join_read_status(h, st, &hs);
    if (rc != RW_OK) return rc;
    own.n_rows += hs.n_store;
    own.keys_upper = hs.n_keys[S];
    rc = join_check_err(h, hs, st);
    if (rc != RW_OK) return rc;
    *out_rows = produced;
  } else {
    JoinScratch sc;
    sc.sortkey = h->sk.as<uint64_t>();
    sc.sortkey_alt = h->sk_alt.as<uint64_t>();
    sc.packed = h->packed.as<uint64_t>();
    sc.offs = h->offs.as<uint64_t>();
    sc.match_slot = h->mslot.as<int64_t>();
    sc.gtable = h->gtable.as<int32_t>();
    sc.gcap = h->gcap;
    uint64_t g = 1024;
    while (g < (uint64_t)n * 2) g <<= 1;
    sc.gcap = g;
    fill_i32_kernel<<<jgrid((int64_t)g, 256), 256, 0, st>>>(sc.gtable, g, -1);
    join_prepare_kernel<<<jgrid(n, 256), 256, 0, st>>>(pd, S, ch, side_dev(h, 1 - S), sc);
    RW_CUDA(cudaGetLastError());
    size_t tb = h->cub_bytes;
    cub::DeviceScan::ExclusiveSum(h->cub_tmp.p, tb, sc.packed, sc.offs, (int)n, st);
    join_totals_kernel<<<1, 1, 0, st>>>(sc.packed, sc.offs, n, ds);
    cub::DoubleBuffer<uint64_t> db(sc.sortkey, sc.sortkey_alt);
    tb = h->cub_bytes;
    cub::DeviceRadixSort::SortKeys(h->cub_tmp.p, tb, db, (int)n, 0, 64, st);
    RW_CUDA(cudaGetLastError());
    h->launches += 6;
    rc = join_read_status(h, st, &hs);
    if (rc != RW_OK) return rc;
    const int64_t reserved = (int64_t)hs.out_rows;
    rc = join_ensure_out(h, reserved);
    if (rc != RW_OK) return rc;
    h->prof.begin(st);
    join_serial_kernel<<<jgrid(n, 128), 128, 0, st>>>(pd, S, ch, side_dev(h, S), side_dev(h, 1 - S), sc, db.Current(), out_dev(h), ds,
                                                        (uint32_t)own.n_rows);
    h->prof.end(st);
    RW_CUDA(cudaGetLastError());
    h->launches++;
    rc = join_read_status(h, st, &hs);
    if (rc != RW_OK) return rc;
    own.n_rows += hs.n_store;
    own.keys_upper = hs.n_keys[S];
    rc = join_check_err(h, hs, st);
    if (rc != RW_OK) return rc;
    *out_rows = reserved;
  }
  RW_CUDA(cudaMemcpyAsync(h->status_host.as<uint8_t>() + 128, h->out_hasnull.p, sizeof(unsigned int) * (J_MAX_OUT + 1), cudaMemcpyDeviceToHost, st));
  RW_CUDA(cudaStreamSynchronize(st));
  memcpy(has_null_host, h->status_host.as<uint8_t>() + 128, sizeof(unsigned int) * (J_MAX_OUT + 1));
  return RW_OK;
}

extern "C" {

int32_t rwgpu_join_create(const rw_join_desc* d, rwgpu_join** out) {
  if (!d || !out) return fail(RW_ERR_INVALID, "null descriptor");
  int rc = rwgpu_device_check();
  if (rc != RW_OK) return rc;
  if (d->join_type < 0 || d->join_type > RW_JOIN_RIGHT_ANTI) return fail(RW_ERR_INVALID, "join type");
  if (d->n_keys < 1 || d->n_keys > RW_MAX_KEYS) return fail(RW_ERR_UNSUPPORTED, "1..4 join key columns supported");
  if (d->left.n_cols > RW_MAX_COLS || d->right.n_cols > RW_MAX_COLS) return fail(RW_ERR_UNSUPPORTED, "too many columns");
  auto h = new rwgpu_join();
  std::unique_ptr<rwgpu_join> guard(h);
  JoinPlanDev& p = h->plan;
  memset(&p, 0, sizeof(p));
  p.T = d->join_type;
  p.n_keys = d->n_keys;
  const rw_join_side_desc* sd[2] = {&d->left, &d->right};
  bool pk_in_jk[2];
  for (int s = 0; s < 2; s++) {
    JoinSideHost& hs = h->side[s];
    hs.n_cols = sd[s]->n_cols;
    hs.types.assign(sd[s]->types, sd[s]->types + sd[s]->n_cols);
    memset(hs.has_valid, 0, sizeof(hs.has_valid));
    p.n_cols[s] = sd[s]->n_cols;
    for (int c = 0; c < sd[s]->n_cols; c++) {
      int w = type_width(sd[s]->types[c]);
      if (!w) return fail(RW_ERR_UNSUPPORTED, "unsupported column type");
      p.col_type[s][c] = sd[s]->types[c];
      p.col_width[s][c] = w;
    }
    for (int k = 0; k < d->n_keys; k++) {
      int c = sd[s]->key_indices[k];
      if (c < 0 || c >= sd[s]->n_cols) return fail(RW_ERR_INVALID, "join key index");
      if (sd[s]->types[c] == RW_T_DECIMAL) return fail(RW_ERR_UNSUPPORTED, "decimal join key");
      p.key_col[s][k] = c;
    }
    p.n_pk[s] = sd[s]->n_pk;
    for (int i = 0; i < sd[s]->n_pk; i++) {
      if (sd[s]->pk_indices[i] < 0 || sd[s]->pk_indices[i] >= sd[s]->n_cols) return fail(RW_ERR_INVALID, "pk index");
      p.pk_col[s][i] = sd[s]->pk_indices[i];
    }
    // pk_contained_in_jk (hash_join.rs:377-378)
    pk_in_jk[s] = true;
    for (int i = 0; i < sd[s]->n_stream_key; i++) {
      bool f = false;
      for (int k = 0; k < d->n_keys; k++) f = f || (sd[s]->key_indices[k] == sd[s]->stream_key[i]);
      pk_in_jk[s] = pk_in_jk[s] && f;
    }
  }
  for (int k = 0; k < d->n_keys; k++) {
    if (d->left.types[d->left.key_indices[k]] != d->right.types[d->right.key_indices[k]])
      return fail(RW_ERR_INVALID, "join key types differ");
    p.null_safe[k] = d->null_safe ? d->null_safe[k] : 0;
  }
  const int T = p.T;
  p.append_only_optimize = d->is_append_only && pk_in_jk[0] && pk_in_jk[1];  // :381
  const bool need_l = (T == RW_JOIN_FULL_OUTER || T == RW_JOIN_LEFT_OUTER || T == RW_JOIN_LEFT_ANTI || T == RW_JOIN_LEFT_SEMI);
  const bool need_r = (T == RW_JOIN_FULL_OUTER || T == RW_JOIN_RIGHT_OUTER || T == RW_JOIN_RIGHT_ANTI || T == RW_JOIN_RIGHT_SEMI);
  p.need_degree[0] = need_l && !pk_in_jk[1];  // :397
  p.need_degree[1] = need_r && !pk_in_jk[0];  // :398
  h->side[0].need_degree = p.need_degree[0];
  h->side[1].need_degree = p.need_degree[1];
  // output schema (:337-359) and i2o mappings (builder.rs:63-80)
  int left_len = d->left.n_cols, right_len = d->right.n_cols;
  std::vector<int> nat;
  if (T == RW_JOIN_LEFT_SEMI || T == RW_JOIN_LEFT_ANTI) { nat.assign(d->left.types, d->left.types + left_len); right_len = 0; }
  else if (T == RW_JOIN_RIGHT_SEMI || T == RW_JOIN_RIGHT_ANTI) { nat.assign(d->right.types, d->right.types + right_len); left_len = 0; }
  else { nat.assign(d->left.types, d->left.types + left_len); nat.insert(nat.end(), d->right.types, d->right.types + right_len); }
  if (d->n_output < 0 || d->n_output > J_MAX_OUT) return fail(RW_ERR_UNSUPPORTED, "too many output columns");
  p.n_out = d->n_output;
  for (int oi = 0; oi < d->n_output; oi++) {
    int idx = d->output_indices[oi];
    if (idx < 0 || idx >= (int)nat.size()) return fail(RW_ERR_INVALID, "output_indices out of bound");
    p.out_type[oi] = nat[idx];
    p.out_width[oi] = type_width(nat[idx]);
    h->out_types.push_back(nat[idx]);
    int s = idx < left_len ? 0 : 1;
    int local = idx < left_len ? idx : idx - left_len;
    p.map_in[s][p.n_map[s]] = local;
    p.map_out[s][p.n_map[s]] = oi;
    p.n_map[s]++;
  }
  p.cond_cmp = d->cond.cmp;
  p.cond_lhs = d->cond.lhs;
  p.cond_rhs = d->cond.rhs;
  if (p.cond_cmp != RW_CMP_NONE) {
    int tot = d->left.n_cols + d->right.n_cols;
    if (p.cond_cmp < 0 || p.cond_cmp > RW_CMP_NE || p.cond_lhs < 0 || p.cond_lhs >= tot || p.cond_rhs < 0 || p.cond_rhs >= tot)
      return fail(RW_ERR_INVALID, "join condition");
    for (int idx : {p.cond_lhs, p.cond_rhs}) {
      int t = idx < d->left.n_cols ? d->left.types[idx] : d->right.types[idx - d->left.n_cols];
      if (type_is_float(t) || t == RW_T_DECIMAL) return fail(RW_ERR_UNSUPPORTED, "non-integer join condition stays on the CPU executor");
    }
  }
  p.single_key = (p.n_keys == 1);
  p.KW = p.single_key ? 1 : 1 + p.n_keys;
  p.SW = p.KW + 1;
  p.strict = d->strict_consistency;
  h->chunk_size = std::max(d->chunk_size > 0 ? d->chunk_size : 1024, 2);  // builder.rs:44-47
  h->fast_inner = (T == RW_JOIN_INNER) && !p.append_only_optimize;

  RW_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  RW_CUDA(h->plan_dev.reserve(sizeof(JoinPlanDev)));
  RW_CUDA(cudaMemcpyAsync(h->plan_dev.p, &p, sizeof(p), cudaMemcpyHostToDevice, h->stream));
  RW_CUDA(h->status.reserve(sizeof(JoinStatus)));
  RW_CUDA(cudaMemsetAsync(h->status.p, 0, sizeof(JoinStatus), h->stream));
  RW_CUDA(h->status_host.reserve(1024));
  RW_CUDA(h->out_hasnull.reserve(sizeof(unsigned int) * (J_MAX_OUT + 1)));
  for (int s = 0; s < 2; s++) {
    uint64_t hint = sd[s]->row_capacity_hint;
    uint64_t cap = 1024;
    while (cap < hint * 2) cap <<= 1;
    h->side[s].slot_cap = cap;
    rc = join_alloc_slots(h, h->side[s].slots, cap);
    if (rc != RW_OK) return rc;
    rc = join_grow_store(h, s, std::max<uint64_t>(hint, 1024));
    if (rc != RW_OK) return rc;
  }
  RW_CUDA(cudaStreamSynchronize(h->stream));
  *out = guard.release();
  return RW_OK;
}

void rwgpu_join_destroy(rwgpu_join* h) {
  if (!h) return;
  if (h->stream) cudaStreamSynchronize(h->stream);
  delete h;
}

int32_t rwgpu_join_push_device(rwgpu_join* h, int32_t side, const rw_chunk* c, rw_chunk* view, void* cuda_stream) {
  if (!h || !c || !view) return fail(RW_ERR_INVALID, "null");
  if (side != 0 && side != 1) return fail(RW_ERR_INVALID, "side");
  if (c->n_cols != h->side[side].n_cols) return fail(RW_ERR_INVALID, "chunk schema mismatch");
  DevChunk ch;
  int rc = devchunk_from_abi(c, &ch);
  if (rc != RW_OK) return rc;
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  int64_t n = 0;
  unsigned int has_null[J_MAX_OUT + 1];
  rc = join_push_dev(h, side, ch, st, &n, has_null);
  if (rc != RW_OK) return rc;
  h->dev_view_cols.resize(h->out_types.size());
  for (size_t k = 0; k < h->out_types.size(); k++) {
    rw_column& col = h->dev_view_cols[k];
    col.type = h->out_types[k];
    col.reserved = 0;
    col.data = h->out_col[k].p;
    col.validity = nullptr;
    if (has_null[k] && n > 0) {
      pack_bytes_to_bits_kernel<<<jgrid((n + 63) / 64, 256), 256, 0, st>>>(h->out_valid[k].as<uint8_t>(), h->out_bits[k].as<uint64_t>(), n);
      RW_CUDA(cudaGetLastError());
      col.validity = h->out_bits[k].as<uint64_t>();
    }
  }
  view->n_rows = n;
  view->n_cols = (int32_t)h->out_types.size();
  view->reserved = 0;
  view->ops = h->out_ops.as<uint8_t>();
  view->visibility = nullptr;
  if (has_null[J_MAX_OUT] && n > 0) {
    pack_bytes_to_bits_kernel<<<jgrid((n + 63) / 64, 256), 256, 0, st>>>(h->out_vis.as<uint8_t>(), h->out_visbits.as<uint64_t>(), n);
    RW_CUDA(cudaGetLastError());
    view->visibility = h->out_visbits.as<uint64_t>();
  }
  view->columns = h->dev_view_cols.data();
  return RW_OK;
}

int32_t rwgpu_join_push(rwgpu_join* h, int32_t side, const rw_chunk* c, rwgpu_out** out) {
  if (!h || !c || !out) return fail(RW_ERR_INVALID, "null");
  if (side != 0 && side != 1) return fail(RW_ERR_INVALID, "side");
  if (c->n_cols != h->side[side].n_cols) return fail(RW_ERR_INVALID, "chunk schema mismatch");
  for (int k = 0; k < c->n_cols; k++)
    if (c->columns[k].type != h->side[side].types[k]) return fail(RW_ERR_INVALID, "chunk column type mismatch");
  // stage the chunk through pinned memory, one H2D copy
  const int64_t n = c->n_rows;
  const size_t nw = (size_t)((n + 63) / 64) * 8;
  size_t total = 256 + (size_t)n + 256 + nw;
  for (int k = 0; k < c->n_cols; k++) total += 512 + (size_t)n * type_width(c->columns[k].type) + nw;
  RW_CUDA(h->up.reserve(total));
  RW_CUDA(h->up_host.reserve(total));
  uint8_t* hp = h->up_host.as<uint8_t>();
  uint8_t* dp = h->up.as<uint8_t>();
  size_t off = 0;
  auto put = [&](const void* src, size_t bytes) -> const void* {
    if (!src) return nullptr;
    size_t o = align_up_j(off, 256);
    memcpy(hp + o, src, bytes);
    off = o + bytes;
    return dp + o;
  };
  DevChunk ch;
  memset(&ch, 0, sizeof(ch));
  ch.n = n;
  ch.n_cols = c->n_cols;
  ch.ops = (const uint8_t*)put(c->ops, (size_t)n);
  ch.vis_bits = (const uint64_t*)put(c->visibility, nw);
  for (int k = 0; k < c->n_cols; k++) {
    int w = type_width(c->columns[k].type);
    ch.cols[k].type = c->columns[k].type;
    ch.cols[k].width = w;
    ch.cols[k].data = put(c->columns[k].data, (size_t)n * w);
    ch.cols[k].valid_bits = (const uint64_t*)put(c->columns[k].validity, nw);
  }
  if (off) RW_CUDA(cudaMemcpyAsync(dp, hp, off, cudaMemcpyHostToDevice, h->stream));
  int64_t rows = 0;
  unsigned int has_null[J_MAX_OUT + 1];
  int rc = join_push_dev(h, side, ch, h->stream, &rows, has_null);
  if (rc != RW_OK) return rc;
  auto o = new rwgpu_out();
  o->n_rows = rows;
  o->chunk_size = h->chunk_size;
  o->types = h->out_types;
  o->ops.resize((size_t)rows);
  o->data.resize(h->out_types.size());
  o->valid_bytes.resize(h->out_types.size());
  if (rows > 0) {
    cudaMemcpyAsync(o->ops.data(), h->out_ops.p, (size_t)rows, cudaMemcpyDeviceToHost, h->stream);
    if (has_null[J_MAX_OUT]) {
      o->vis_bytes.resize((size_t)rows);
      cudaMemcpyAsync(o->vis_bytes.data(), h->out_vis.p, (size_t)rows, cudaMemcpyDeviceToHost, h->stream);
    }
    for (size_t k = 0; k < h->out_types.size(); k++) {
      size_t w = type_width(h->out_types[k]);
      o->data[k].resize((size_t)rows * w);
      cudaMemcpyAsync(o->data[k].data(), h->out_col[k].p, (size_t)rows * w, cudaMemcpyDeviceToHost, h->stream);
      if (has_null[k]) {
        o->valid_bytes[k].resize((size_t)rows);
        cudaMemcpyAsync(o->valid_bytes[k].data(), h->out_valid[k].p, (size_t)rows, cudaMemcpyDeviceToHost, h->stream);
      }
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { delete o; return fail(RW_ERR_CUDA, cudaGetErrorString(e)); }
  }
  o->finalize();
  *out = o;
  return RW_OK;
}

int32_t rwgpu_join_barrier(rwgpu_join* h, uint64_t /*epoch*/) {
  if (!h) return fail(RW_ERR_INVALID, "null");
  // state lives in HBM (StateStore stubbed to memory, north_star): a barrier is an ordering point
  RW_CUDA(cudaStreamSynchronize(h->stream));
  return RW_OK;
}

int32_t rwgpu_join_profile(rwgpu_join* h, int32_t enable, double* ms, uint64_t* launches) {
  if (!h) return fail(RW_ERR_INVALID, "null");
  RW_CUDA(cudaDeviceSynchronize());
  h->prof.collect();
  if (ms) *ms = h->prof.ms;
  if (launches) *launches = h->prof.n;
  h->prof.ms = 0;
  h->prof.n = 0;
  h->prof.on = enable != 0;
  return RW_OK;
}

int32_t rwgpu_join_stats(rwgpu_join* h, uint64_t* left_rows, uint64_t* right_rows, uint64_t* launches) {
  if (!h) return fail(RW_ERR_INVALID, "null");
  JoinStatus s;
  RW_CUDA(cudaStreamSynchronize(h->stream));
  RW_CUDA(cudaMemcpy(&s, h->status.p, sizeof(s), cudaMemcpyDeviceToHost));
  if (left_rows) *left_rows = h->side[0].n_rows;
  if (right_rows) *right_rows = h->side[1].n_rows;
  if (launches) *launches = h->launches;
  return RW_OK;
}

}  // extern "C"
