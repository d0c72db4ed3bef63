// join.cu -- placeholder until the join kernels land (returns RW_ERR_UNSUPPORTED).
#include "common.cuh"
using namespace rw;
struct rwgpu_join { int dummy; };
extern "C" {
int32_t rwgpu_join_create(const rw_join_desc*, rwgpu_join**) { return fail(RW_ERR_UNSUPPORTED, "join not built yet"); }
void rwgpu_join_destroy(rwgpu_join*) {}
int32_t rwgpu_join_push(rwgpu_join*, int32_t, const rw_chunk*, rwgpu_out**) { return fail(RW_ERR_UNSUPPORTED, "join"); }
int32_t rwgpu_join_push_device(rwgpu_join*, int32_t, const rw_chunk*, rw_chunk*, void*) { return fail(RW_ERR_UNSUPPORTED, "join"); }
int32_t rwgpu_join_barrier(rwgpu_join*, uint64_t) { return fail(RW_ERR_UNSUPPORTED, "join"); }
int32_t rwgpu_join_stats(rwgpu_join*, uint64_t*, uint64_t*, uint64_t*) { return fail(RW_ERR_UNSUPPORTED, "join"); }
}
