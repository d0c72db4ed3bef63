// CPU-only check of the C++ host mirror's BufferedWatermarks (include/rwgpu_executor.hpp) against the traces of
// src/stream/src/executor/hash_join.rs:3648-3715 (test_streaming_hash_join_watermark) and a three-upstream case.
#include <cstdio>

#include "rwgpu_executor.hpp"

using rwgpu::BufferedWatermarks;
using rwgpu::Watermark;

static int fails = 0;
#define EXPECT(c) do { if (!(c)) { printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); fails++; } } while (0)

int main() {
  {
    BufferedWatermarks b({0, 1});  // left, right
    EXPECT(!b.handle_watermark(0, Watermark{0, 4, 100}));
    EXPECT(!b.handle_watermark(0, Watermark{0, 4, 200}));
    auto w = b.handle_watermark(1, Watermark{0, 4, 50});
    EXPECT(w && w->val == 50);
    w = b.handle_watermark(1, Watermark{0, 4, 100});
    EXPECT(w && w->val == 100);
    w = b.handle_watermark(1, Watermark{0, 4, 300});
    EXPECT(w && w->val == 200);
  }
  {
    BufferedWatermarks b({0, 1, 2});
    EXPECT(!b.handle_watermark(0, Watermark{0, 4, 7}));
    EXPECT(!b.handle_watermark(1, Watermark{0, 4, 5}));
    EXPECT(!b.handle_watermark(1, Watermark{0, 4, 9}));
    auto w = b.handle_watermark(2, Watermark{0, 4, 5});
    EXPECT(w && w->val == 5);
    w = b.handle_watermark(2, Watermark{0, 4, 8});
    EXPECT(w && w->val == 7);
  }
  printf(fails ? "watermarks: %d FAILED\n" : "watermarks: ok\n", fails);
  return fails ? 1 : 0;
}
