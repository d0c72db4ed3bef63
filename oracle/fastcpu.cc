// fastcpu.cc -- performance-minded CPU restatement of the two benchmarked operator shapes, used ONLY
// as the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`).
//
// *** TEST / MEASUREMENT INFRASTRUCTURE ONLY (same rule as oracle.cc). ***
//
// Why it exists: the reference is Rust and cannot be built here, and oracle.cc is written for
// clarity (std::map, per-datum boxed values), which would flatter the GPU.  This file keeps the
// reference's *structure* -- one single-threaded actor per partition (actor.rs:272), row-at-a-time
// probing (hash_join.rs:977), per-key entry sets (join/hash_join.rs:736-830), per-datum output
// append through a chunk builder cut at 1024 rows (join/builder.rs:84-148), per-group value states
// with barrier-time change inference (agg_group.rs:131-166,431-606) -- but with flat open-addressed
// tables and typed int64 columns, and with the StateTable mem-table write elided, i.e. every choice
// favours the CPU.  tests/test_fastcpu.py checks its results against oracle.cc.
//
// Shapes: join  = Inner, key = column 0 (int64), pk = column 1, 4 int64 columns per side, all 8
//                 output columns (BASELINE cfg3, Nexmark q7/q8-shaped bid x auction);
//         agg   = group key column 0 (int64), count(*), sum(col1)->int8, max(col1) append-only or
//                 count/sum with retractions (BASELINE cfg2, Nexmark q4-shaped).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

struct JRow { int64_t c[4]; };

// per-key entry set: up to 2 rows inline, then a heap vector (JoinRowSet keeps a Vec up to 4 rows)
struct Entry {
  int64_t key;
  uint32_t n;
  uint32_t used;
  JRow inl[2];
  std::vector<JRow>* more;
};

struct SideTable {
  std::vector<Entry> slots;
  uint64_t mask = 0, count = 0;
  SideTable() { resize(1 << 16); }
  void resize(uint64_t cap) {
    std::vector<Entry> old;
    old.swap(slots);
    slots.assign(cap, Entry{0, 0, 0, {}, nullptr});
    mask = cap - 1;
    count = 0;
    for (auto& e : old)
      if (e.used) { Entry* d = find_or_insert(e.key); *d = e; }
  }
  Entry* find(int64_t key) {
    uint64_t i = mix64((uint64_t)key) & mask;
    while (slots[i].used) {
      if (slots[i].key == key) return &slots[i];
      i = (i + 1) & mask;
    }
    return nullptr;
  }
  Entry* find_or_insert(int64_t key) {
    if ((count + 1) * 2 > slots.size()) resize(slots.size() * 2);
    uint64_t i = mix64((uint64_t)key) & mask;
    while (slots[i].used) {
      if (slots[i].key == key) return &slots[i];
      i = (i + 1) & mask;
    }
    slots[i].used = 1;
    slots[i].key = key;
    slots[i].n = 0;
    slots[i].more = nullptr;
    count++;
    return &slots[i];
  }
};

inline JRow* entry_row(Entry* e, uint32_t i) { return i < 2 ? &e->inl[i] : &(*e->more)[i - 2]; }

struct OutBuilder {  // StreamChunkBuilder with 8 int64 columns, cut at 1024 rows
  std::vector<int64_t> col[8];
  std::vector<uint8_t> ops;
  uint64_t total = 0, chunks = 0;
  uint64_t checksum = 0;
  OutBuilder() { for (auto& c : col) c.reserve(1024); ops.reserve(1024); }
  inline void append(uint8_t op, const int64_t* u, const int64_t* m, bool u_left) {
    ops.push_back(op);
    const int64_t* l = u_left ? u : m;
    const int64_t* r = u_left ? m : u;
    for (int k = 0; k < 4; k++) col[k].push_back(l[k]);
    for (int k = 0; k < 4; k++) col[4 + k].push_back(r[k]);
    // order-independent checksum of the emitted (op, row) multiset: every column enters with its own odd weight
    static const uint64_t W[8] = {3, 31, 5, 7, 11, 1, 17, 19};
    uint64_t v = 0;
    for (int k = 0; k < 4; k++) v += W[k] * (uint64_t)l[k] + W[4 + k] * (uint64_t)r[k];
    checksum += v * (op == 1 ? 1 : (uint64_t)-1);
    total++;
    if (ops.size() == 1024) take();
  }
  void take() {
    if (ops.empty()) return;
    chunks++;
    for (auto& c : col) c.clear();
    ops.clear();
  }
};

}  // namespace

struct rwf_join {
  SideTable side[2];
  OutBuilder out;
};

struct AggGroup {
  int64_t key;
  int64_t cnt, sum, mx;
  int64_t p_cnt, p_sum, p_mx;
  uint8_t used, dirty, has_prev, sum_some;
};

struct rwf_agg {
  std::vector<AggGroup> slots;
  std::vector<uint32_t> dirty;
  uint64_t mask = 0, count = 0;
  int with_max = 1;
  uint64_t out_rows = 0, checksum = 0;
  rwf_agg() { resize(1 << 16); }
  void resize(uint64_t cap) {
    std::vector<AggGroup> old;
    old.swap(slots);
    slots.assign(cap, AggGroup{});
    mask = cap - 1;
    count = 0;
    std::vector<uint32_t> nd;
    for (auto& g : old)
      if (g.used) {
        AggGroup* d = find_or_insert(g.key);
        *d = g;
        if (g.dirty) nd.push_back((uint32_t)(d - slots.data()));
      }
    dirty.swap(nd);
  }
  AggGroup* find_or_insert(int64_t key) {
    if ((count + 1) * 2 > slots.size()) resize(slots.size() * 2);
    uint64_t i = mix64((uint64_t)key) & mask;
    while (slots[i].used) {
      if (slots[i].key == key) return &slots[i];
      i = (i + 1) & mask;
    }
    AggGroup& g = slots[i];
    g = AggGroup{};
    g.used = 1;
    g.key = key;
    g.mx = INT64_MIN;
    count++;
    return &g;
  }
};

extern "C" {

rwf_join* rwf_join_new() { return new rwf_join(); }
void rwf_join_free(rwf_join* h) {
  for (auto& s : h->side)
    for (auto& e : s.slots) delete e.more;
  delete h;
}

// eq_join_oneside for one chunk of `n` rows (ops 1..4, 4 int64 columns); returns emitted rows
int64_t rwf_join_push(rwf_join* h, int side, int64_t n, const uint8_t* ops, const int64_t* c0, const int64_t* c1,
                      const int64_t* c2, const int64_t* c3) {
  SideTable& own = h->side[side];
  SideTable& other = h->side[1 - side];
  const uint64_t before = h->out.total;
  for (int64_t r = 0; r < n; r++) {
    const int64_t u[4] = {c0[r], c1[r], c2[r], c3[r]};
    const bool ins = (ops[r] == 1 || ops[r] == 3);
    if (Entry* e = other.find(u[0])) {
      for (uint32_t i = 0; i < e->n; i++) h->out.append(ins ? 1 : 2, u, entry_row(e, i)->c, side == 0);
    }
    if (ins) {
      Entry* e = own.find_or_insert(u[0]);
      JRow row;
      memcpy(row.c, u, sizeof(u));
      if (e->n < 2) e->inl[e->n] = row;
      else {
        if (!e->more) e->more = new std::vector<JRow>();
        e->more->push_back(row);
      }
      e->n++;
    } else if (Entry* e = own.find(u[0])) {
      for (uint32_t i = 0; i < e->n; i++) {
        if (entry_row(e, i)->c[1] == u[1]) {  // remove by pk: swap_remove (join_row_set.rs:103-107)
          *entry_row(e, i) = *entry_row(e, e->n - 1);
          if (e->n > 2) e->more->pop_back();
          e->n--;
          break;
        }
      }
    }
  }
  h->out.take();  // final partial chunk (hash_join.rs:1059-1061)
  return (int64_t)(h->out.total - before);
}
uint64_t rwf_join_checksum(rwf_join* h) { return h->out.checksum; }
uint64_t rwf_join_out_rows(rwf_join* h) { return h->out.total; }

rwf_agg* rwf_agg_new(int with_max) {
  auto* h = new rwf_agg();
  h->with_max = with_max;
  return h;
}
void rwf_agg_free(rwf_agg* h) { delete h; }

// apply_chunk: row-at-a-time state update (hash_agg.rs:332-409 without the per-group bitmaps)
void rwf_agg_push(rwf_agg* h, int64_t n, const uint8_t* ops, const int64_t* key, const int64_t* val) {
  for (int64_t r = 0; r < n; r++) {
    AggGroup* g = h->find_or_insert(key[r]);
    if (!g->dirty) { g->dirty = 1; h->dirty.push_back((uint32_t)(g - h->slots.data())); }
    const bool retract = (ops[r] == 2 || ops[r] == 4);
    g->cnt += retract ? -1 : 1;
    g->sum_some = 1;
    int64_t s;
    if (retract ? __builtin_sub_overflow(g->sum, val[r], &s) : __builtin_add_overflow(g->sum, val[r], &s)) abort();
    g->sum = s;
    if (h->with_max && val[r] > g->mx) g->mx = val[r];
  }
}

// flush_data at a barrier: change inference per dirty group, output rows appended (and counted)
int64_t rwf_agg_flush(rwf_agg* h) {
  uint64_t before = h->out_rows;
  for (uint32_t gi : h->dirty) {
    AggGroup& g = h->slots[gi];
    g.dirty = 0;
    if (g.cnt == 0) { g.sum = 0; g.mx = INT64_MIN; g.sum_some = 0; }
    const int64_t prc = g.has_prev ? g.p_cnt : 0;
    if (prc == 0 && g.cnt == 0) continue;
    if (prc == 0) { h->out_rows += 1; h->checksum += (uint64_t)(g.key + g.cnt + g.sum + (h->with_max ? g.mx : 0)); }
    else if (g.cnt == 0) { h->out_rows += 1; h->checksum -= (uint64_t)(g.key + g.p_cnt + g.p_sum + (h->with_max ? g.p_mx : 0)); g.has_prev = 0; continue; }
    else if (g.p_cnt != g.cnt || g.p_sum != g.sum || (h->with_max && g.p_mx != g.mx)) {
      h->out_rows += 2;
      h->checksum -= (uint64_t)(g.key + g.p_cnt + g.p_sum + (h->with_max ? g.p_mx : 0));
      h->checksum += (uint64_t)(g.key + g.cnt + g.sum + (h->with_max ? g.mx : 0));
    } else continue;
    g.p_cnt = g.cnt; g.p_sum = g.sum; g.p_mx = g.mx; g.has_prev = 1;
  }
  h->dirty.clear();
  return (int64_t)(h->out_rows - before);
}
uint64_t rwf_agg_checksum(rwf_agg* h) { return h->checksum; }
uint64_t rwf_agg_groups(rwf_agg* h) { return h->count; }

}  // extern "C"

// pre-size a side's table for `n` keys (avoids rehash storms while loading the build side)
extern "C" void rwf_join_reserve(rwf_join* h, int side, uint64_t n) {
  uint64_t cap = 1 << 16;
  while (cap < n * 2) cap <<= 1;
  if (cap > h->side[side].slots.size()) h->side[side].resize(cap);
}
extern "C" void rwf_agg_reserve(rwf_agg* h, uint64_t n) {
  uint64_t cap = 1 << 16;
  while (cap < n * 2) cap <<= 1;
  if (cap > h->slots.size()) h->resize(cap);
}

#include <thread>
// Actor pool: P long-lived OS threads, one actor each (the reference runs one long-lived task per actor on a
// worker pool, actor.rs:209-232,272).  What round 1 got wrong and this version fixes:
//   * worker a is pinned to cpu_ids[a], a list the caller takes from sched_getaffinity (one entry per
//     PHYSICAL core), not to `a % nprocs` -- CPU ids of an allowed set need not be 0..P-1; the return code of
//     the pin is kept and reported;
//   * every actor and its tables are created and first-touched ON ITS OWN THREAD (NUMA first touch);
//   * actors do not meet at a condition variable after every 2^20-row step: each consumes ITS stream of
//     batches independently, exactly as the reference's actors do between barriers; one spin barrier after the
//     warm-up batches opens the timed region, wall time = last finish - first start, per-actor busy time is
//     reported so the caller can state the parallel efficiency.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>
struct rwf_pool {
  int P = 0;
  std::vector<rwf_join*> actors;
  std::vector<std::thread> th;
  std::vector<int> cpu, pin_rc;
  std::mutex m;
  std::condition_variable cv_start, cv_done;
  uint64_t gen = 0;
  int remaining = 0;
  bool stop = false;
  // task of the current generation
  int kind = 0;  // 0 = reserve, 1 = run
  const uint64_t* reserve[2] = {nullptr, nullptr};
  int side = 0, chunk = 1024, nb = 0, warm = 0;
  const int64_t* n = nullptr;                 // [nb][P]
  const uint8_t* const* ops = nullptr;        // [nb][P]
  const int64_t* const* c[4] = {nullptr, nullptr, nullptr, nullptr};
  std::atomic<int> at_barrier{0};
  std::vector<int64_t> outs;
  std::vector<double> t0, t1;
};
static inline double rwf_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void rwf_pool_worker(rwf_pool* p, int a) {
  if (p->cpu[a] >= 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(p->cpu[a], &set);
    p->pin_rc[a] = pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
  }
  p->actors[a] = new rwf_join();  // first touch on this thread
  uint64_t seen = 0;
  {
    std::lock_guard<std::mutex> lk(p->m);
    if (--p->remaining == 0) p->cv_done.notify_one();
  }
  while (true) {
    {
      std::unique_lock<std::mutex> lk(p->m);
      p->cv_start.wait(lk, [&] { return p->stop || p->gen != seen; });
      if (p->stop) return;
      seen = p->gen;
    }
    if (p->kind == 0) {
      for (int s = 0; s < 2; s++)
        if (p->reserve[s]) rwf_join_reserve(p->actors[a], s, p->reserve[s][a]);
    } else {
      int64_t tot = 0;
      auto run_batch = [&](int b) {
        const int64_t na = p->n[(size_t)b * p->P + a];
        const size_t k = (size_t)b * p->P + a;
        for (int64_t i = 0; i < na; i += p->chunk) {
          const int64_t m = na - i < p->chunk ? na - i : p->chunk;
          tot += rwf_join_push(p->actors[a], p->side, m, p->ops[k] + i, p->c[0][k] + i, p->c[1][k] + i, p->c[2][k] + i, p->c[3][k] + i);
        }
      };
      for (int b = 0; b < p->warm; b++) run_batch(b);
      p->at_barrier.fetch_add(1);
      while (p->at_barrier.load(std::memory_order_acquire) < p->P) sched_yield();
      p->t0[a] = rwf_now();
      tot = 0;
      for (int b = p->warm; b < p->nb; b++) run_batch(b);
      p->t1[a] = rwf_now();
      p->outs[a] = tot;
    }
    {
      std::lock_guard<std::mutex> lk(p->m);
      if (--p->remaining == 0) p->cv_done.notify_one();
    }
  }
}
static void rwf_pool_dispatch(rwf_pool* p) {
  {
    std::lock_guard<std::mutex> lk(p->m);
    p->remaining = p->P;
    p->gen++;
  }
  p->cv_start.notify_all();
  std::unique_lock<std::mutex> lk(p->m);
  p->cv_done.wait(lk, [&] { return p->remaining == 0; });
}
// cpu_ids: P CPU ids (-1 = do not pin) or NULL (no pinning)
extern "C" rwf_pool* rwf_pool_new(int P, const int* cpu_ids) {
  rwf_pool* p = new rwf_pool();
  p->P = P;
  p->actors.assign(P, nullptr);
  p->cpu.assign(P, -1);
  p->pin_rc.assign(P, 0);
  if (cpu_ids) p->cpu.assign(cpu_ids, cpu_ids + P);
  p->outs.assign(P, 0);
  p->t0.assign(P, 0);
  p->t1.assign(P, 0);
  p->remaining = P;
  for (int a = 0; a < P; a++) p->th.emplace_back(rwf_pool_worker, p, a);
  std::unique_lock<std::mutex> lk(p->m);
  p->cv_done.wait(lk, [&] { return p->remaining == 0; });
  return p;
}
extern "C" int rwf_pool_pin_failures(rwf_pool* p) {
  int f = 0;
  for (int a = 0; a < p->P; a++) f += p->cpu[a] >= 0 && p->pin_rc[a] != 0;
  return f;
}
extern "C" rwf_join* rwf_pool_actor(rwf_pool* p, int a) { return p->actors[a]; }
// table sizing per actor, done by the actor's own thread (keys_left / keys_right: P entries each, may be NULL)
extern "C" void rwf_pool_reserve(rwf_pool* p, const uint64_t* keys_left, const uint64_t* keys_right) {
  p->kind = 0;
  p->reserve[0] = keys_left;
  p->reserve[1] = keys_right;
  rwf_pool_dispatch(p);
}
// every actor consumes its `nb` batches of side `side` (n / ops / c0..c3 indexed [batch * P + actor]) in `chunk`-row
// StreamChunks; the first `warm` batches are untimed.  *wall_s = last finish - first start of the timed part,
// busy_s[P] = per-actor time in the timed part.  Returns the rows emitted in the timed part.
extern "C" int64_t rwf_pool_run(rwf_pool* p, int side, int nb, int warm, const int64_t* n, const uint8_t* const* ops,
                                const int64_t* const* c0, const int64_t* const* c1, const int64_t* const* c2,
                                const int64_t* const* c3, int chunk, double* wall_s, double* busy_s) {
  p->kind = 1;
  p->side = side; p->nb = nb; p->warm = warm; p->n = n; p->ops = ops; p->chunk = chunk;
  p->c[0] = c0; p->c[1] = c1; p->c[2] = c2; p->c[3] = c3;
  p->at_barrier.store(0);
  rwf_pool_dispatch(p);
  double first = p->t0[0], last = p->t1[0];
  int64_t tot = 0;
  for (int a = 0; a < p->P; a++) {
    first = p->t0[a] < first ? p->t0[a] : first;
    last = p->t1[a] > last ? p->t1[a] : last;
    if (busy_s) busy_s[a] = p->t1[a] - p->t0[a];
    tot += p->outs[a];
  }
  if (wall_s) *wall_s = last - first;
  return tot;
}
extern "C" uint64_t rwf_pool_checksum(rwf_pool* p) {
  uint64_t s = 0;
  for (auto* a : p->actors) s += a->out.checksum;
  return s;
}
extern "C" void rwf_pool_free(rwf_pool* p) {
  {
    std::lock_guard<std::mutex> lk(p->m);
    p->stop = true;
  }
  p->cv_start.notify_all();
  for (auto& t : p->th) t.join();
  for (auto* a : p->actors) if (a) rwf_join_free(a);
  delete p;
}
