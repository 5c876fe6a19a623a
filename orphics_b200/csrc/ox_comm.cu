// The one exchange step of the path: an NCCL sum all-reduce over NVLink of the Statistics triple
// (stats.py:1215-1217) and of the mean-field stack + count (stats.py:1227-1228), driven from the C ABI so that
// any host -- not only Python -- can reduce.  One process per GPU; the host distributes the 128-byte NCCL
// unique id through whatever rendezvous it has (torch.distributed's store in bench.py, MPI_Bcast, a file).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): liborphx.so keeps loading on hosts without NCCL, and a
// process that already holds torch's bundled NCCL shares that copy instead of loading a second one.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include "ox_common.cuh"

using namespace ox;

struct ox_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1, device = 0;
};

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.lib) return OX_OK;
  const char *names[] = {getenv("ORPHX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    if (!n || !n[0]) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("NCCL not found (dlopen libnccl.so.2 failed: %s); set ORPHX_NCCL_LIB", dlerror());
    return OX_ERR_UNSUPPORTED;
  }
#define OX_SYM(field, name)                                                   \
  do {                                                                        \
    *(void **)(&g_nccl.field) = dlsym(h, name);                               \
    if (!g_nccl.field) {                                                      \
      set_error("NCCL symbol %s missing", name);                              \
      return OX_ERR_UNSUPPORTED;                                              \
    }                                                                         \
  } while (0)
  OX_SYM(GetUniqueId, "ncclGetUniqueId");
  OX_SYM(CommInitRank, "ncclCommInitRank");
  OX_SYM(CommDestroy, "ncclCommDestroy");
  OX_SYM(AllReduce, "ncclAllReduce");
  OX_SYM(GetErrorString, "ncclGetErrorString");
  OX_SYM(GetVersion, "ncclGetVersion");
#undef OX_SYM
  g_nccl.lib = h;
  return OX_OK;
}

#define OX_NCCL(call)                                                                              \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess) {                                                                       \
      set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(r_));         \
      return OX_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

}  // namespace

namespace ox {
int comm_allreduce_f64(ox_comm *c, double *buf, long long count) {
  OX_REQUIRE(c && buf && count >= 0, "all-reduce: null pointer");
  if (c->nranks == 1 || count == 0) return OX_OK;
  OX_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, c->comm, g_stream));
  g_launches++;
  return OX_OK;
}
}  // namespace ox

extern "C" {

int ox_comm_unique_id(void *id, size_t len) {
  OX_REQUIRE(id && len >= sizeof(ncclUniqueId), "ox_comm_unique_id: the buffer must hold %zu bytes", sizeof(ncclUniqueId));
  OX_TRY(load_nccl());
  ncclUniqueId u;
  OX_NCCL(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return OX_OK;
}

int ox_comm_create(int rank, int nranks, const void *id, size_t len, ox_comm **out) {
  OX_REQUIRE(out && nranks >= 1 && rank >= 0 && rank < nranks, "ox_comm_create: rank %d of %d", rank, nranks);
  ox_comm *c = new ox_comm;
  c->rank = rank;
  c->nranks = nranks;
  cudaGetDevice(&c->device);
  if (nranks > 1) {
    if (!id || len < sizeof(ncclUniqueId)) {
      delete c;
      set_error("ox_comm_create: the unique id of rank 0 (ox_comm_unique_id, %zu bytes) is required", sizeof(ncclUniqueId));
      return OX_ERR_INVALID;
    }
    int st = load_nccl();
    if (st != OX_OK) { delete c; return st; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) {
      set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
      delete c;
      return OX_ERR_CUDA;
    }
  }
  *out = c;
  return OX_OK;
}

int ox_comm_destroy(ox_comm *c) {
  if (!c) return OX_OK;
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  return OX_OK;
}

int ox_comm_info(ox_comm *c, int *rank, int *nranks, int *nccl_version) {
  OX_REQUIRE(c, "null communicator");
  if (rank) *rank = c->rank;
  if (nranks) *nranks = c->nranks;
  if (nccl_version) {
    *nccl_version = 0;
    if (g_nccl.lib) g_nccl.GetVersion(nccl_version);
  }
  return OX_OK;
}

int ox_comm_allreduce_f64(ox_comm *c, double *buf_dev, long long count) { return comm_allreduce_f64(c, buf_dev, count); }

int ox_pipeline_allreduce(ox_comm *c, ox_pipeline *pl) {
  OX_REQUIRE(c && pl, "null pointer");
  const long long d = pl->dim;
  return comm_allreduce_f64(c, pl->stat.as<double>(), 1 + d + d * d);   // [N | SUM | CROSS] in one collective
}

}  // extern "C"
