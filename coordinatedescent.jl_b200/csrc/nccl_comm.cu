// nccl_comm.cu — the one collective of the path: summing row-sharded partial Gram matrices
// (G | c in one buffer) over NVLink with ncclAllReduce(sum, f64).  One process per GPU; the
// 128-byte ncclUniqueId is produced by rank 0 (cdgpu_comm_unique_id) and distributed by the host
// side (torch.distributed / MPI / Julia's Distributed).  NCCL is resolved with dlopen at first use
// so libcdgpu.so loads on machines without it (and shares torch's copy when torch loaded it first).
#include <dlfcn.h>

#include "common.cuh"

#define API extern "C" __attribute__((visibility("default")))

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int load_nccl() {
  if (g_nccl.lib) return CDGPU_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *nm : names) {
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return cdgpu_set_error(CDGPU_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(lib, "ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
    return cdgpu_set_error(CDGPU_ENCCL, "libnccl is missing required symbols");
  g_nccl.lib = lib;
  return CDGPU_OK;
}
int nccl_fail(const char *what, ncclResult_t r) {
  return cdgpu_set_error(CDGPU_ENCCL, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
}
} // namespace

struct cdgpu_comm_s {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1, device = 0;
};

API int cdgpu_comm_unique_id(void *id128) {
  return api_guard([&]() -> int {
  if (!id128) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  CD_TRY(load_nccl());
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r) return nccl_fail("ncclGetUniqueId", r);
  memcpy(id128, &id, sizeof id);
  return CDGPU_OK;
  });
}

API int cdgpu_comm_init(cdgpu_comm *c, const void *id128, int rank, int nranks, int device) {
  return api_guard([&]() -> int {
  if (!c || !id128) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (nranks < 1 || rank < 0 || rank >= nranks) return cdgpu_set_error(CDGPU_EARG, "bad rank / nranks");
  CD_TRY(load_nccl());
  CUDA_TRY(cudaSetDevice(device));
  cdgpu_comm_s *cc = new cdgpu_comm_s();
  cc->rank = rank;
  cc->nranks = nranks;
  cc->device = device;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclResult_t r = g_nccl.CommInitRank(&cc->comm, nranks, id, rank);
  if (r) {
    delete cc;
    return nccl_fail("ncclCommInitRank", r);
  }
  *c = cc;
  return CDGPU_OK;
  });
}

API int cdgpu_comm_destroy(cdgpu_comm c) {
  if (!c) return CDGPU_OK;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
  return CDGPU_OK;
}

int cdgpu_comm_allreduce(cdgpu_comm c, double *buf, size_t count, cudaStream_t s) {
  if (!c || c->nranks == 1) return CDGPU_OK;
  ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, c->comm, s);
  if (r) return nccl_fail("ncclAllReduce", r);
  return CDGPU_OK;
}
