// e2e_caller.cc -- the `e2e` leg of bench.py as a COMPILED caller of the C ABI (what the Rust shim does per message):
// pinned host StreamChunk buffers -> rwgpu_join_push -> every output chunk view is fetched and read -> release.
// (Round 1 timed this from Python and touched only the first and the last of the 1024 output chunk views per step,
// because walking all of them through ctypes cost a millisecond of interpreter time per step.)
//
// Workload = bench.py's headline (BASELINE configs[2]): 10 M auction rows loaded on the right, then steps of 2^20 bid
// rows (1024 chunks of 1024 rows coalesced into one call) on the left; same splitmix64 generators as bench.py.
// Output: one JSON object on stdout.
// argv[5] = "async": the shim's pipelined form -- rwgpu_join_push_async of step s+1 is issued before the output of step s
// is collected (rwgpu_join_collect_out), so input H2D and output D2H of neighbouring steps overlap.
// Build: g++ -O2 -std=c++17 -Iinclude tools/e2e_caller.cc -o build/e2e_caller -Lrisingwave_b200 -lrwgpu -L/usr/local/cuda/lib64 -lcudart
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "rwgpu.h"

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static const uint64_t SEED = 0x20210410;

struct Pinned {
  void* p = nullptr;
  explicit Pinned(size_t bytes) { if (cudaMallocHost(&p, bytes) != cudaSuccess) { fprintf(stderr, "cudaMallocHost failed\n"); exit(2); } }
  ~Pinned() { cudaFreeHost(p); }
};

#define CHECK(rc)                                                                              \
  do {                                                                                         \
    int rc_ = (rc);                                                                            \
    if (rc_ != RW_OK) { fprintf(stderr, "rwgpu status %d: %s\n", rc_, rwgpu_last_error()); exit(3); } \
  } while (0)

int main(int argc, char** argv) {
  const int64_t n_build = argc > 1 ? atoll(argv[1]) : 10000000, batch = argc > 2 ? atoll(argv[2]) : (1 << 20);
  const int steps = argc > 3 ? atoi(argv[3]) : 20, warmup = argc > 4 ? atoi(argv[4]) : 3;
  const bool async = argc > 5 && !strcmp(argv[5], "async");
  const int32_t t4[4] = {RW_T_INT64, RW_T_INT64, RW_T_INT64, RW_T_INT64};
  const int32_t key0[1] = {0}, pk_l[1] = {1}, sk_l[1] = {1}, sk_r[1] = {0};
  const uint8_t null_safe[1] = {0};
  const int32_t outs[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  rw_join_desc d;
  memset(&d, 0, sizeof(d));
  d.join_type = RW_JOIN_INNER;
  d.n_keys = 1;
  d.left = {4, t4, key0, 1, pk_l, 1, sk_l, (uint64_t)n_build};
  d.right = {4, t4, key0, 0, pk_l, 1, sk_r, (uint64_t)n_build};
  d.null_safe = null_safe;
  d.n_output = 8;
  d.output_indices = outs;
  d.chunk_size = 1024;
  d.strict_consistency = 1;
  rwgpu_join* h = nullptr;
  CHECK(rwgpu_join_create(&d, &h));

  int aliased = 0;  // output columns whose views point into the caller's input buffers (rwgpu.h: not copied back)
  // what the shim does with a result: every chunk view, one value of it read
  auto consume = [&](rwgpu_out* out, int64_t* const cols[4], int64_t n, int64_t* rows_out, uint64_t* touched) {
    const int nch = rwgpu_out_num_chunks(out);
    for (int i = 0; i < nch; i++) {
      rw_chunk v;
      CHECK(rwgpu_out_chunk(out, i, &v));
      if (v.n_rows) *touched += (uint64_t)((const int64_t*)v.columns[v.n_cols - 1].data)[v.n_rows - 1] + v.ops[0];
      if (i == 0 && v.n_rows) {
        aliased = 0;
        for (int k = 0; k < v.n_cols; k++)
          for (int c2 = 0; c2 < 4; c2++)
            if ((const int64_t*)v.columns[k].data >= cols[c2] && (const int64_t*)v.columns[k].data < cols[c2] + n) aliased++;
      }
    }
    *rows_out += rwgpu_out_num_rows(out);
    rwgpu_out_release(out);
  };
  auto push = [&](int side, const uint8_t* ops, int64_t* const cols[4], int64_t n, int64_t* rows_out, uint64_t* touched) {
    rw_column c[4];
    for (int k = 0; k < 4; k++) c[k] = {RW_T_INT64, 0, cols[k], nullptr, nullptr};
    rw_chunk ch = {n, 4, 0, ops, nullptr, c};
    rwgpu_out* out = nullptr;
    CHECK(rwgpu_join_push(h, side, &ch, &out));
    consume(out, cols, n, rows_out, touched);
  };
  auto launch = [&](int side, const uint8_t* ops, int64_t* const cols[4], int64_t n) {
    rw_column c[4];
    for (int k = 0; k < 4; k++) c[k] = {RW_T_INT64, 0, cols[k], nullptr, nullptr};
    rw_chunk ch = {n, 4, 0, ops, nullptr, c};
    CHECK(rwgpu_join_push_async(h, side, &ch));
  };

  // build side: ids in a pseudo-random arrival order (bench.py gen_auctions)
  std::vector<int64_t> pos_of((size_t)n_build);  // auction id -> arrival position (the verification step needs the row of an id)
  {
    std::vector<int64_t> order(n_build);
    std::iota(order.begin(), order.end(), 0);
    std::vector<uint64_t> key(n_build);
    for (int64_t i = 0; i < n_build; i++) key[i] = splitmix64((uint64_t)i ^ SEED);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key[a] < key[b]; });
    for (int64_t g = 0; g < n_build; g++) pos_of[(size_t)order[g]] = g;
    Pinned p((size_t)batch * 33);
    int64_t* cols[4];
    for (int k = 0; k < 4; k++) cols[k] = (int64_t*)p.p + (size_t)k * batch;
    uint8_t* ops = (uint8_t*)((int64_t*)p.p + 4 * (size_t)batch);
    memset(ops, RW_OP_INSERT, (size_t)batch);
    int64_t rows = 0;
    uint64_t t = 0;
    for (int64_t lo = 0; lo < n_build; lo += batch) {
      const int64_t m = std::min(batch, n_build - lo);
      for (int64_t i = 0; i < m; i++) {
        const uint64_t g = (uint64_t)(lo + i);
        cols[0][i] = order[lo + i];
        cols[1][i] = (int64_t)(splitmix64(g ^ (SEED + 1)) % 1000000ull);
        cols[2][i] = 10 + (int64_t)(splitmix64(g ^ (SEED + 2)) % 5ull);
        cols[3][i] = (int64_t)(splitmix64(g ^ (SEED + 3)) % (1ull << 40));
      }
      push(RW_SIDE_RIGHT, ops, cols, m, &rows, &t);
    }
  }
  // probe batches in pinned memory (the shim's StreamChunk arrays live in a pinned arena)
  const int total = warmup + steps;
  std::vector<Pinned*> bufs;
  for (int s = 0; s < total + 1; s++) {  // (+1: the verification step)
    bufs.push_back(new Pinned((size_t)batch * 33));
    int64_t* base = (int64_t*)bufs.back()->p;
    for (int64_t i = 0; i < batch; i++) {
      const uint64_t g = (uint64_t)((int64_t)s * batch + i);
      base[i] = (int64_t)(splitmix64(g ^ (SEED + 10)) % (uint64_t)n_build);
      base[batch + i] = (int64_t)g + 1600000000000000ll;
      base[2 * batch + i] = (int64_t)(splitmix64(g ^ (SEED + 11)) % 1000000ull);
      base[3 * batch + i] = (int64_t)(splitmix64(g ^ (SEED + 12)) % (1ull << 24));
    }
    memset(base + 4 * batch, RW_OP_INSERT, (size_t)batch);
  }
  int64_t rows = 0;
  uint64_t touched = 0;
  auto step = [&](int s) {
    int64_t* base = (int64_t*)bufs[s]->p;
    int64_t* cols[4] = {base, base + batch, base + 2 * batch, base + 3 * batch};
    push(RW_SIDE_LEFT, (const uint8_t*)(base + 4 * batch), cols, batch, &rows, &touched);
  };
  auto step_launch = [&](int s) {
    int64_t* base = (int64_t*)bufs[s]->p;
    int64_t* cols[4] = {base, base + batch, base + 2 * batch, base + 3 * batch};
    launch(RW_SIDE_LEFT, (const uint8_t*)(base + 4 * batch), cols, batch);
  };
  auto step_collect = [&](int s) {
    int64_t* base = (int64_t*)bufs[s]->p;
    int64_t* cols[4] = {base, base + batch, base + 2 * batch, base + 3 * batch};
    rwgpu_out* out = nullptr;
    CHECK(rwgpu_join_collect_out(h, &out));
    consume(out, cols, batch, &rows, &touched);
  };
  // steps lo .. hi-1; async: step s+1 is launched before step s is collected
  auto run = [&](int lo, int hi) {
    if (!async) { for (int s = lo; s < hi; s++) step(s); return; }
    for (int s = lo; s < hi; s++) {
      step_launch(s);
      if (s > lo) step_collect(s - 1);
    }
    step_collect(hi - 1);
  };
  run(0, warmup);
  cudaDeviceSynchronize();
  rows = 0;
  const auto t0 = std::chrono::steady_clock::now();
  run(warmup, total);
  cudaDeviceSynchronize();
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const int64_t timed_rows = rows;
  // ---- verification (untimed): one more step through the same path, EVERY output row read; row count and the
  // order-independent checksum  sum over rows of sign(op) * sum_k w_k * col_k  (bench.py CHECKSUM_WEIGHTS) are compared
  // with the join evaluated directly on the host (bid row x the auction row of its id)
  unsigned long long cs_got = 0, cs_want = 0;
  long long rows_got = 0, rows_want = 0;
  {
    static const unsigned long long W8[8] = {3, 31, 5, 7, 11, 1, 17, 19};
    int64_t* base = (int64_t*)bufs[total]->p;
    int64_t* cols[4] = {base, base + batch, base + 2 * batch, base + 3 * batch};
    rw_column c[4];
    for (int k = 0; k < 4; k++) c[k] = {RW_T_INT64, 0, cols[k], nullptr, nullptr};
    rw_chunk ch = {batch, 4, 0, (const uint8_t*)(base + 4 * batch), nullptr, c};
    rwgpu_out* out = nullptr;
    if (async) { CHECK(rwgpu_join_push_async(h, RW_SIDE_LEFT, &ch)); CHECK(rwgpu_join_collect_out(h, &out)); }
    else CHECK(rwgpu_join_push(h, RW_SIDE_LEFT, &ch, &out));
    const int nch = rwgpu_out_num_chunks(out);
    for (int i = 0; i < nch; i++) {
      rw_chunk v;
      CHECK(rwgpu_out_chunk(out, i, &v));
      for (int64_t r = 0; r < v.n_rows; r++) {
        if (v.visibility && !((v.visibility[r >> 6] >> (r & 63)) & 1ull)) continue;
        unsigned long long x = 0;
        for (int k = 0; k < 8; k++) x += W8[k] * (unsigned long long)((const int64_t*)v.columns[k].data)[r];
        const bool ins = v.ops[r] == RW_OP_INSERT || v.ops[r] == RW_OP_UPDATE_INSERT;
        cs_got += ins ? x : 0ull - x;
        rows_got++;
      }
    }
    rwgpu_out_release(out);
    for (int64_t i = 0; i < batch; i++) {
      const int64_t id = cols[0][i];
      const uint64_t g = (uint64_t)pos_of[(size_t)id];
      const int64_t a[4] = {id, (int64_t)(splitmix64(g ^ (SEED + 1)) % 1000000ull), 10 + (int64_t)(splitmix64(g ^ (SEED + 2)) % 5ull),
                            (int64_t)(splitmix64(g ^ (SEED + 3)) % (1ull << 40))};
      unsigned long long x = 0;
      for (int k = 0; k < 4; k++) x += W8[k] * (unsigned long long)cols[k][i] + W8[4 + k] * (unsigned long long)a[k];
      cs_want += x;
      rows_want++;
    }
  }
  rows = timed_rows;
  printf("{\"value\": %.1f, \"unit\": \"rows/s\", \"ms_per_step\": %.4f, \"steps\": %d, \"out_rows\": %lld, \"h2d_bytes_per_step\": %lld, "
         "\"chunk_views_read_per_step\": %lld, \"output_columns_aliasing_input\": %d, \"touched\": %llu, \"mode\": \"%s\", "
         "\"verified\": %s, \"verify_rows\": [%lld, %lld], \"verify_checksum\": [\"%016llx\", \"%016llx\"]}\n",
         (double)steps * (double)batch / dt, dt / steps * 1e3, steps, (long long)rows, (long long)(batch * 33), (long long)(batch / 1024), aliased,
         (unsigned long long)touched, async ? "async" : "sync", (rows_got == rows_want && cs_got == cs_want) ? "true" : "false", rows_got, rows_want,
         cs_got, cs_want);
  rwgpu_join_destroy(h);
  for (auto* b : bufs) delete b;
  return 0;
}
