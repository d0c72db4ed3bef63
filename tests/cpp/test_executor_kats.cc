// C++ host-layer test: replays reference golden tests through include/rwgpu_executor.hpp (the C++
// mirror of the executor interface) on the GPU.  Test data transcribed from
//   src/stream/src/executor/hash_join.rs:1812-1880  test_streaming_hash_inner_join
//   src/stream/src/executor/hash_join.rs:3347-3430  test_streaming_hash_full_outer_join
//   src/stream/tests/integration_tests/hash_agg.rs:21-96  test_hash_agg_count_sum
//   src/stream/src/executor/filter.rs:208-272  test_filter  (exact: ops and visibility of every row)
// Comparison is the net applied multiset per expected chunk (the reference's output order is not
// deterministic).  Exit code 0 = all passed.
#include <cstdio>
#include <iostream>

#include "rwgpu_executor.hpp"

using namespace rwgpu;

static int failures = 0;
#define EXPECT(cond, msg) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, msg); failures++; } } while (0)

static StreamChunk next_chunk(Execute& ex) {
  auto m = ex.poll_next();
  if (!m || !std::holds_alternative<StreamChunk>(*m)) { std::printf("FAIL expected a ready chunk\n"); failures++; return StreamChunk(); }
  return std::get<StreamChunk>(*m);
}
static void next_pending(Execute& ex) {
  auto m = ex.poll_next();
  // a chunk whose rows cancel is as good as pending for the applied result
  if (m && std::holds_alternative<StreamChunk>(*m) && net_multiset({std::get<StreamChunk>(*m)}).empty()) m = ex.poll_next();
  EXPECT(!m, "expected pending");
}
static void next_barrier(Execute& ex) {
  auto m = ex.poll_next();
  EXPECT(m && std::holds_alternative<Barrier>(*m), "expected barrier");
}
static void expect_same(const StreamChunk& got, const char* want_pretty) {
  EXPECT(net_multiset({got}) == net_multiset({StreamChunk::from_pretty(want_pretty)}), want_pretty);
}

static void test_streaming_hash_inner_join() {
  const std::vector<int32_t> T = {RW_T_INT64, RW_T_INT64};
  auto tx_l = std::make_shared<MockSource>(T, std::vector<int32_t>{1});
  auto tx_r = std::make_shared<MockSource>(T, std::vector<int32_t>{1});
  HashJoinExecutor hash_join(RW_JOIN_INNER, tx_l, tx_r, JoinParams{{0}, {1}}, JoinParams{{0}, {1}}, {0});
  tx_l->push_barrier(1); tx_r->push_barrier(1); next_barrier(hash_join);
  tx_l->push_chunk(StreamChunk::from_pretty(" I I\n + 1 4\n + 2 5\n + 3 6")); next_pending(hash_join);
  tx_l->push_barrier(2); tx_r->push_barrier(2); next_barrier(hash_join);
  tx_l->push_chunk(StreamChunk::from_pretty(" I I\n + 3 8\n - 3 8")); next_pending(hash_join);
  tx_r->push_chunk(StreamChunk::from_pretty(" I I\n + 2 7\n + 4 8\n + 6 9"));
  expect_same(next_chunk(hash_join), " I I I I\n + 2 5 2 7");
  tx_r->push_chunk(StreamChunk::from_pretty(" I I\n + 3 10\n + 6 11"));
  expect_same(next_chunk(hash_join), " I I I I\n + 3 6 3 10");
}

static void test_streaming_hash_full_outer_join() {
  const std::vector<int32_t> T = {RW_T_INT64, RW_T_INT64};
  auto tx_l = std::make_shared<MockSource>(T, std::vector<int32_t>{1});
  auto tx_r = std::make_shared<MockSource>(T, std::vector<int32_t>{1});
  HashJoinExecutor hash_join(RW_JOIN_FULL_OUTER, tx_l, tx_r, JoinParams{{0}, {1}}, JoinParams{{0}, {1}}, {0});
  tx_l->push_barrier(1); tx_r->push_barrier(1); next_barrier(hash_join);
  tx_l->push_chunk(StreamChunk::from_pretty(" I I\n + 1 4\n + 2 5\n + 3 6"));
  expect_same(next_chunk(hash_join), " I I I I\n + 1 4 . .\n + 2 5 . .\n + 3 6 . .");
  tx_l->push_chunk(StreamChunk::from_pretty(" I I\n + 3 8\n - 3 8"));
  expect_same(next_chunk(hash_join), " I I I I\n + 3 8 . .\n - 3 8 . .");
  tx_r->push_chunk(StreamChunk::from_pretty(" I I\n + 2 7\n + 4 8\n + 6 9"));
  expect_same(next_chunk(hash_join), " I I I I\n - 2 5 . .\n + 2 5 2 7\n + . . 4 8\n + . . 6 9");
  tx_r->push_chunk(StreamChunk::from_pretty(" I I\n + 3 10\n + 6 11"));
  expect_same(next_chunk(hash_join), " I I I I\n - 3 6 . .\n + 3 6 3 10\n + . . 6 11");
}

static void test_hash_agg_count_sum() {
  auto tx = std::make_shared<MockSource>(std::vector<int32_t>{RW_T_INT64, RW_T_INT64, RW_T_INT64}, std::vector<int32_t>{});
  HashAggExecutor hash_agg(tx, false, {{RW_AGG_COUNT, -1, RW_T_INT64}, {RW_AGG_SUM, 1, RW_T_INT64}, {RW_AGG_SUM, 2, RW_T_INT64}}, 0, {0});
  tx->push_barrier(1);
  tx->push_chunk(StreamChunk::from_pretty(" I I I\n + 1 1 1\n + 2 2 2\n + 2 2 2"));
  tx->push_barrier(2);
  tx->push_chunk(StreamChunk::from_pretty(" I I I\n - 1 1 1\n - 2 2 2 D\n - 2 2 2\n + 3 3 3"));
  tx->push_barrier(3);
  next_barrier(hash_agg);
  expect_same(next_chunk(hash_agg), " I I I I\n + 1 1 1 1\n + 2 2 4 4");
  next_barrier(hash_agg);
  expect_same(next_chunk(hash_agg), " I I I I\n + 3 1 3 3\n - 1 1 1 1\n U- 2 2 4 4\n U+ 2 1 2 2");
  next_barrier(hash_agg);
  // a plan the device path does not take: create must refuse so the shim falls back to the CPU executor
  try {
    HashAggExecutor bad(tx, false, {{RW_AGG_COUNT, -1, RW_T_INT64}, {RW_AGG_SUM, 1, RW_T_DECIMAL}}, 0, {0, 1, 2, 0, 1});  // 5 group key columns
    EXPECT(false, "more than 4 group key columns must be unsupported");
  } catch (const StreamExecutorError& e) {
    EXPECT(e.code == RW_ERR_UNSUPPORTED, "error code");
  }
  // retractable min = MaterializedInput state (hash_agg.rs test_hash_agg_min, tests/integration_tests/hash_agg.rs:97-170)
  {
    auto tm = std::make_shared<MockSource>(std::vector<int32_t>{RW_T_INT64, RW_T_INT64, RW_T_INT64}, std::vector<int32_t>{});
    HashAggExecutor agg_min(tm, false, {{RW_AGG_COUNT, -1, RW_T_INT64}, {RW_AGG_MIN, 1, RW_T_INT64}}, 0, {0});
    tm->push_barrier(1);
    tm->push_chunk(StreamChunk::from_pretty(" I I I\n + 1 233 1001\n + 1 23333 1002\n + 2 2333 1003"));
    tm->push_barrier(2);
    tm->push_chunk(StreamChunk::from_pretty(" I I I\n - 1 233 1001\n - 1 23333 1002 D\n - 2 2333 1003"));
    tm->push_barrier(3);
    next_barrier(agg_min);
    expect_same(next_chunk(agg_min), " I I I\n + 1 2 233\n + 2 1 2333");
    next_barrier(agg_min);
    expect_same(next_chunk(agg_min), " I I I\n - 2 1 2333\n U- 1 2 233\n U+ 1 1 23333");
    next_barrier(agg_min);
  }
}

static void expect_exact(const StreamChunk& got, const char* want_pretty) {
  const StreamChunk want = StreamChunk::from_pretty(want_pretty);
  bool same = got.capacity() == want.capacity() && got.ops == want.ops;
  for (int64_t r = 0; same && r < got.capacity(); r++) same = got.is_visible(r) == want.is_visible(r);
  EXPECT(same, want_pretty);
}

static void test_filter() {
  auto tx = std::make_shared<MockSource>(std::vector<int32_t>{RW_T_INT64, RW_T_INT64}, std::vector<int32_t>{});
  // (greater_than:boolean $0:int8 $1:int8)
  FilterExecutor filter(tx, {rw_filter_term{RW_CMP_GT, 0, 1, 0, 0}});
  tx->push_chunk(StreamChunk::from_pretty(" I I\n + 1 4\n + 5 2\n + 6 6\n - 7 5"));
  tx->push_chunk(StreamChunk::from_pretty(" I I\n U- 5 3\n U+ 7 5\n U- 5 3\n U+ 3 5\n U- 3 5\n U+ 5 3\n U- 3 5\n U+ 4 6"));
  expect_exact(next_chunk(filter), " I I\n + 1 4 D\n + 5 2\n + 6 6 D\n - 7 5");
  expect_exact(next_chunk(filter), " I I\n U- 5 3\n U+ 7 5\n - 5 3\n U+ 3 5 D\n U- 3 5 D\n + 5 3\n U- 3 5 D\n U+ 4 6 D");
}

int main() {
  if (rwgpu_device_check() != RW_OK) { std::printf("no CUDA device: %s\n", rwgpu_last_error()); return 2; }
  test_streaming_hash_inner_join();
  test_streaming_hash_full_outer_join();
  test_hash_agg_count_sum();
  test_filter();
  std::printf(failures ? "%d FAILED\n" : "all C++ host-layer KATs passed (%d failures)\n", failures);
  return failures ? 1 : 0;
}
