// emu_lib_stubs.cpp — the entry points of sharded.cu (NCCL + CUDA IPC: nothing the CPU emulation can stand in for)
// answer COZO_GPU_EUNSUP in libcozo_gpu_emu.so, so that the library still exports every symbol of include/cozo_gpu.h.
#include "common.cuh"

using namespace cozo;
static int unsup() { return set_error(COZO_GPU_EUNSUP, "the sharded operator is not part of the CPU emulation build"); }
struct cozo_gpu_shards {};

extern "C" {
int cozo_gpu_shards_unique_id(uint8_t*) { return unsup(); }
int cozo_gpu_shards_init(cozo_gpu_shards_t** out, const uint8_t*, int, int) {
  if (out) *out = nullptr;
  return unsup();
}
void cozo_gpu_shards_free(cozo_gpu_shards_t*) {}
int cozo_gpu_shards_info(cozo_gpu_shards_t*, int*, int*, int*, uint64_t*) { return unsup(); }
int cozo_gpu_hnsw_stage_sharded(cozo_gpu_shards_t*, cozo_gpu_hnsw_t*, uint64_t*, uint64_t*) { return unsup(); }
int cozo_gpu_hnsw_search_sharded(cozo_gpu_shards_t*, const float*, uint32_t, uint32_t, uint32_t, double, int, uint64_t*, float*,
                                 uint32_t*, CozoGpuSearchStats*) { return unsup(); }
int cozo_gpu_hnsw_search_sharded_dev(cozo_gpu_shards_t*, const float*, uint32_t, uint32_t, uint32_t, double, uint64_t*, float*,
                                     uint32_t*, void*) { return unsup(); }
}
