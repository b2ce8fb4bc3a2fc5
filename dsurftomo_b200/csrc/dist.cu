// Multi-GPU plumbing for the distributed LSMR: NCCL is resolved at run time with dlopen (the
// library that torch.distributed already loaded in the process, or the wheel-bundled
// libnccl.so.2), so libdsurf_b200.so itself has no link-time NCCL dependency and loads on
// machines without it.  One process per GPU; the host side (dsurftomo_b200/dist.py) creates
// the ncclUniqueId on rank 0 and broadcasts it through torch.distributed.
//
// Exchange step of the path: rows of A are partitioned over ranks, the
// n-vectors are replicated; each iteration needs exactly one all-reduce of the partial
// A'u (n floats) fused (ncclGroup) with the partial ||u||^2 (one double).
#include <dlfcn.h>
#include <cstring>
#include "../../include/dsurftomo_b200.h"
#include "common.cuh"

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7, ncclFloat64 = 8 };  // ncclDataType_t
enum { ncclSum = 0 };                       // ncclRedOp_t
typedef int (*fn_GetUniqueId)(ncclUniqueId *);
typedef int (*fn_CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
typedef int (*fn_Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_Group)(void);
typedef const char *(*fn_ErrStr)(int);
struct Nccl {
  void *h = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_AllReduce AllReduce = nullptr;
  fn_AllGather AllGather = nullptr;
  fn_Broadcast Broadcast = nullptr;
  fn_Group GroupStart = nullptr, GroupEnd = nullptr;
  fn_ErrStr ErrStr = nullptr;
} g;

int load_nccl() {
  if (g.h) return DSURF_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  void *h = nullptr;
  for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    if (const char *p = getenv("DSURF_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!h) {
    dsurf::set_error(__FILE__, __LINE__, "cannot dlopen libnccl.so.2 (import torch first or set DSURF_NCCL_LIB)");
    return DSURF_ERR_NCCL;
  }
  g.GetUniqueId = (fn_GetUniqueId)dlsym(h, "ncclGetUniqueId");
  g.CommInitRank = (fn_CommInitRank)dlsym(h, "ncclCommInitRank");
  g.CommDestroy = (fn_CommDestroy)dlsym(h, "ncclCommDestroy");
  g.AllReduce = (fn_AllReduce)dlsym(h, "ncclAllReduce");
  g.AllGather = (fn_AllGather)dlsym(h, "ncclAllGather");
  g.Broadcast = (fn_Broadcast)dlsym(h, "ncclBroadcast");
  g.GroupStart = (fn_Group)dlsym(h, "ncclGroupStart");
  g.GroupEnd = (fn_Group)dlsym(h, "ncclGroupEnd");
  g.ErrStr = (fn_ErrStr)dlsym(h, "ncclGetErrorString");
  if (!g.GetUniqueId || !g.CommInitRank || !g.CommDestroy || !g.AllReduce || !g.AllGather || !g.Broadcast ||
      !g.GroupStart || !g.GroupEnd) {
    dsurf::set_error(__FILE__, __LINE__, "libnccl lacks an expected symbol");
    return DSURF_ERR_NCCL;
  }
  g.h = h;
  return DSURF_OK;
}
int nccl_check(int rc, int line) {
  if (rc == ncclSuccess) return DSURF_OK;
  dsurf::set_error(__FILE__, line, g.ErrStr ? g.ErrStr(rc) : "nccl error");
  return DSURF_ERR_NCCL;
}
}  // namespace

namespace dsurf {
// the group is always closed, also when a call inside it failed (an open group would poison every later NCCL call)
int lsmr_allreduce(void *comm, float *buf, size_t n, double *dbuf, size_t nd, cudaStream_t st) {
  if (!comm) return DSURF_OK;
  DS_CHECK(load_nccl());
  DS_CHECK(nccl_check(g.GroupStart(), __LINE__));
  int rc = DSURF_OK;
  if (n) rc = nccl_check(g.AllReduce(buf, buf, n, ncclFloat32, ncclSum, (ncclComm_t)comm, st), __LINE__);
  if (nd && rc == DSURF_OK) rc = nccl_check(g.AllReduce(dbuf, dbuf, nd, ncclFloat64, ncclSum, (ncclComm_t)comm, st), __LINE__);
  const int rc2 = nccl_check(g.GroupEnd(), __LINE__);
  return rc != DSURF_OK ? rc : rc2;
}

// ---- exchange of the forward/sensitivity path: every rank contributes one block
int nccl_allgather_bytes(void *comm, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st) {
  DS_CHECK(load_nccl());
  return nccl_check(g.AllGather(send, recv, bytes_per_rank, /*ncclInt8*/ 0, (ncclComm_t)comm, st), __LINE__);
}
// grouped broadcasts: block r (count4[r] 32-bit words at recv + off4[r]) comes from rank r; `send` is this rank's block
int nccl_allgatherv_words(void *comm, int rank, int nranks, const void *send, void *recv, const long long *off4,
                          const long long *count4, cudaStream_t st) {
  DS_CHECK(load_nccl());
  DS_CHECK(nccl_check(g.GroupStart(), __LINE__));
  int rc = DSURF_OK;
  for (int r = 0; r < nranks && rc == DSURF_OK; r++) {
    if (count4[r] <= 0) continue;
    char *dst = (char *)recv + 4 * off4[r];
    rc = nccl_check(g.Broadcast(r == rank ? send : (const void *)dst, dst, (size_t)count4[r], ncclFloat32, r, (ncclComm_t)comm, st),
                    __LINE__);
  }
  const int rc2 = nccl_check(g.GroupEnd(), __LINE__);
  return rc != DSURF_OK ? rc : rc2;
}
}  // namespace dsurf

extern "C" int dsurf_nccl_unique_id(void *id128) {
  DS_CHECK(load_nccl());
  ncclUniqueId id;
  DS_CHECK(nccl_check(g.GetUniqueId(&id), __LINE__));
  memcpy(id128, &id, 128);
  return DSURF_OK;
}
extern "C" int dsurf_nccl_comm_init(void **comm, const void *id128, int rank, int nranks) {
  DS_CHECK(dsurf::ensure_device());
  DS_CHECK(load_nccl());
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  DS_CHECK(nccl_check(g.CommInitRank(&c, nranks, id, rank), __LINE__));
  *comm = c;
  return DSURF_OK;
}
extern "C" int dsurf_nccl_comm_destroy(void *comm) {
  if (!comm) return DSURF_OK;
  DS_CHECK(load_nccl());
  return nccl_check(g.CommDestroy((ncclComm_t)comm), __LINE__);
}
