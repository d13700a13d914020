// comm.cu -- the one exchange step of the path inside the library: NCCL all-gather of projective partial sums.
//
// an MSM sharded by point range needs a single gather of xyzz points (160 - 480 B per rank and MSM;
// all five MSMs of a proof in one exchange).  A Rust / C++ caller drives it through the C ABI alone:
//   rank 0: pcdgpu_comm_unique_id(id) -> ships the 128 bytes to the other ranks by any means (MPI, a file, a pipe);
//   every rank: pcdgpu_comm_init(ctx, id, rank, world) -> pcdgpu_pk_upload_sharded / pcdgpu_groth16_prove_sharded,
//               pcdgpu_msm_bases_sharded.
// libnccl is resolved at run time (dlopen of the process's libnccl.so.2: the one torch already loaded under torchrun,
// else the system's), so libpcdgpu.so has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return &api;
#define PCD_NCCL_SYM(field, sym) *(void**)(&api.field) = dlsym(api.handle, sym)
  PCD_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  PCD_NCCL_SYM(CommInitRank, "ncclCommInitRank");
  PCD_NCCL_SYM(CommDestroy, "ncclCommDestroy");
  PCD_NCCL_SYM(AllGather, "ncclAllGather");
  PCD_NCCL_SYM(GroupStart, "ncclGroupStart");
  PCD_NCCL_SYM(GroupEnd, "ncclGroupEnd");
  PCD_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef PCD_NCCL_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GroupStart && api.GroupEnd;
  return &api;
}
int nccl_fail(pcdgpu_ctx* ctx, const char* what, ncclResult_t r) {
  NcclApi* a = nccl_api();
  if (ctx) ctx->set_error("%s: %s", what, a->GetErrorString ? a->GetErrorString(r) : "NCCL error");
  return PCDGPU_E_CUDA;
}
}  // namespace

// internal interface (groth16.cuh): all-gather `bytes` bytes per rank from d_send into d_recv (rank-major) on `st`
int comm_allgather(pcdgpu_ctx* ctx, const void* d_send, void* d_recv, size_t bytes, cudaStream_t st) {
  if (ctx->comm_world == 1) {
    if (d_send != d_recv) PCD_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  NcclApi* a = nccl_api();
  ncclResult_t r = a->AllGather(d_send, d_recv, bytes, ncclChar, (ncclComm_t)ctx->nccl_comm, st);
  if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllGather", r);
  return 0;
}
int comm_group(pcdgpu_ctx* ctx, bool start) {
  if (ctx->comm_world == 1) return 0;
  NcclApi* a = nccl_api();
  ncclResult_t r = start ? a->GroupStart() : a->GroupEnd();
  if (r != ncclSuccess) return nccl_fail(ctx, start ? "ncclGroupStart" : "ncclGroupEnd", r);
  return 0;
}

extern "C" {

int pcdgpu_comm_unique_id(void* out_id) {
  if (!out_id) return PCDGPU_E_ARG;
  NcclApi* a = nccl_api();
  if (!a->ok) return PCDGPU_E_NODEVICE;
  ncclUniqueId id;
  if (a->GetUniqueId(&id) != ncclSuccess) return PCDGPU_E_CUDA;
  static_assert(sizeof(id) == PCDGPU_COMM_ID_BYTES, "ncclUniqueId size");
  memcpy(out_id, &id, sizeof(id));
  return 0;
}

int pcdgpu_comm_init(pcdgpu_ctx* ctx, const void* id, int rank, int world) {
  if (!ctx) return PCDGPU_E_ARG;
  if (world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) {
    ctx->set_error("bad rank / world size / id");
    return PCDGPU_E_ARG;
  }
  if (ctx->nccl_comm || ctx->comm_world != 1) {
    ctx->set_error("the context already has a communicator");
    return PCDGPU_E_ARG;
  }
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  if (world > 1) {
    NcclApi* a = nccl_api();
    if (!a->ok) {
      ctx->set_error("libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbols");
      return PCDGPU_E_NODEVICE;
    }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm;
    ncclResult_t r = a->CommInitRank(&comm, world, uid, rank);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->nccl_comm = comm;
  }
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return 0;
}

int pcdgpu_comm_info(const pcdgpu_ctx* ctx, int* rank, int* world) {
  if (!ctx) return PCDGPU_E_ARG;
  if (rank) *rank = ctx->comm_rank;
  if (world) *world = ctx->comm_world;
  return 0;
}

void pcdgpu_comm_destroy(pcdgpu_ctx* ctx) {
  if (!ctx) return;
  if (ctx->nccl_comm) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl_api()->CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
}

}  // extern "C"
