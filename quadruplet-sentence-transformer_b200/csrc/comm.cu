// Collectives of the corpus-sharded path behind the C ABI: NCCL over NVLink 5 /
// NVSwitch called DIRECTLY, so a C caller (or the Python mirror without torch.distributed) can run the
// sharded retrieval.  libnccl.so.2 is resolved at run time with dlopen -- the copy already loaded in the
// process (PyTorch bundles one) is reused when there is one -- and only its stable 2.x entry points are
// used, declared here (no nccl.h needed to build).
//
// The reference has no multi-GPU path; what these exchanges replace is the sequential corpus chunk loop
// of /root/reference/ir_evauation_script.py:161 (corpus_chunk_size), cut in space instead of time.
#include "qst_common.cuh"
#include <dlfcn.h>
#include <mutex>

namespace qst {

typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
enum { kNcclInt8 = 0, kNcclFloat32 = 7 };
enum { kNcclMax = 2 };

struct NcclApi {
  int (*GetUniqueId)(nccl_unique_id*);
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int);
  int (*CommDestroy)(nccl_comm_t);
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
  int (*GetVersion)(int*);
  bool ok;
};

static NcclApi g_nccl{};
static std::once_flag g_nccl_once;

static void load_nccl() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy torch (or the caller) already loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return;
#define QST_NCCL_SYM(field, name)                                            \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
  if (!g_nccl.field) return;
  QST_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  QST_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  QST_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  QST_NCCL_SYM(AllGather, "ncclAllGather")
  QST_NCCL_SYM(AllReduce, "ncclAllReduce")
  QST_NCCL_SYM(Send, "ncclSend")
  QST_NCCL_SYM(Recv, "ncclRecv")
  QST_NCCL_SYM(GroupStart, "ncclGroupStart")
  QST_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  QST_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  QST_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef QST_NCCL_SYM
  g_nccl.ok = true;
}

static const NcclApi* nccl() {
  std::call_once(g_nccl_once, load_nccl);
  return g_nccl.ok ? &g_nccl : nullptr;
}

#define QST_NCCL(call)                                                                             \
  do {                                                                                             \
    int r__ = (call);                                                                              \
    if (r__ != 0) {                                                                                \
      ::qst::set_error("%s failed: %s (%s:%d)", #call, api->GetErrorString(r__), __FILE__, __LINE__); \
      return QST_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

}  // namespace qst

using namespace qst;

struct qst_comm {
  nccl_comm_t comm;
  int world, rank;
};

extern "C" int qst_comm_available(void) { return nccl() != nullptr ? 1 : 0; }

extern "C" int qst_comm_nccl_version(void) {
  const NcclApi* api = nccl();
  int v = 0;
  if (!api || api->GetVersion(&v) != 0) return 0;
  return v;
}

extern "C" int qst_comm_unique_id(unsigned char* id128) {
  QST_CHECK_ARG(id128 != nullptr, "comm_unique_id: null argument");
  const NcclApi* api = nccl();
  if (!api) { set_error("libnccl.so.2 could not be loaded"); return QST_ERR_UNSUPPORTED; }
  nccl_unique_id id;
  QST_NCCL(api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return QST_OK;
}

extern "C" int qst_comm_init(const unsigned char* id128, int world, int rank, qst_comm** out) {
  QST_CHECK_ARG(id128 && out, "comm_init: null argument");
  QST_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "comm_init: bad world=%d rank=%d", world, rank);
  const NcclApi* api = nccl();
  if (!api) { set_error("libnccl.so.2 could not be loaded"); return QST_ERR_UNSUPPORTED; }
  nccl_unique_id id;
  memcpy(&id, id128, sizeof(id));
  qst_comm* c = new qst_comm{nullptr, world, rank};
  int r = api->CommInitRank(&c->comm, world, id, rank);
  if (r != 0) {
    set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
    delete c;
    return QST_ERR_CUDA;
  }
  *out = c;
  return QST_OK;
}

extern "C" int qst_comm_destroy(qst_comm* c) {
  if (!c) return QST_OK;
  const NcclApi* api = nccl();
  if (api && c->comm) api->CommDestroy(c->comm);
  delete c;
  return QST_OK;
}

extern "C" int qst_comm_world(const qst_comm* c) { return c ? c->world : 0; }
extern "C" int qst_comm_rank(const qst_comm* c) { return c ? c->rank : -1; }

extern "C" int qst_comm_allgather(qst_comm* c, const void* send, void* recv, size_t bytes_per_rank, qst_stream_t stream) {
  QST_CHECK_ARG(c && send && recv, "comm_allgather: null argument");
  const NcclApi* api = nccl();
  QST_NCCL(api->AllGather(send, recv, bytes_per_rank, kNcclInt8, c->comm, reinterpret_cast<cudaStream_t>(stream)));
  return QST_OK;
}

// Block r of `send` (bytes_per_peer bytes) goes to rank r; block r of `recv` came from rank r.
extern "C" int qst_comm_alltoall(qst_comm* c, const void* send, void* recv, size_t bytes_per_peer, qst_stream_t stream) {
  QST_CHECK_ARG(c && send && recv, "comm_alltoall: null argument");
  const NcclApi* api = nccl();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const char* s = reinterpret_cast<const char*>(send);
  char* d = reinterpret_cast<char*>(recv);
  QST_NCCL(api->GroupStart());
  for (int r = 0; r < c->world; ++r) {
    int e = api->Send(s + (size_t)r * bytes_per_peer, bytes_per_peer, kNcclInt8, r, c->comm, st);
    if (e == 0) e = api->Recv(d + (size_t)r * bytes_per_peer, bytes_per_peer, kNcclInt8, r, c->comm, st);
    if (e != 0) {
      api->GroupEnd();
      set_error("ncclSend/ncclRecv failed: %s", api->GetErrorString(e));
      return QST_ERR_CUDA;
    }
  }
  QST_NCCL(api->GroupEnd());
  return QST_OK;
}

extern "C" int qst_comm_allreduce_max_f32(qst_comm* c, const float* send, float* recv, size_t n, qst_stream_t stream) {
  QST_CHECK_ARG(c && send && recv, "comm_allreduce_max_f32: null argument");
  const NcclApi* api = nccl();
  QST_NCCL(api->AllReduce(send, recv, n, kNcclFloat32, kNcclMax, c->comm, reinterpret_cast<cudaStream_t>(stream)));
  return QST_OK;
}

// K6's exchange: [Q, k] per rank -> [G, Q, k] on every rank (rank-major), scores and ids.
extern "C" int qst_allgather_topk(qst_comm* c, const float* vals, const int64_t* idx, int64_t Q, int k, float* out_vals,
                                  int64_t* out_idx, qst_stream_t stream) {
  QST_CHECK_ARG(Q >= 0 && k >= 1, "allgather_topk: bad shape Q=%lld k=%d", (long long)Q, k);
  if (Q == 0) return QST_OK;
  int rc = qst_comm_allgather(c, vals, out_vals, (size_t)Q * k * sizeof(float), stream);
  if (rc) return rc;
  return qst_comm_allgather(c, idx, out_idx, (size_t)Q * k * sizeof(int64_t), stream);
}

// The candidate-list exchange of the sharded path: lists [G * q_own, m + 1] of 8-byte entries grouped by
// owner rank (qst_select_candidates over all G * q_own queries) -> recv [G, q_own, m + 1] (source-major).
extern "C" int qst_exchange_candidates(qst_comm* c, const void* lists, void* recv, int64_t q_own, int m,
                                       qst_stream_t stream) {
  QST_CHECK_ARG(q_own >= 1 && m >= 1, "exchange_candidates: bad shape q_own=%lld m=%d", (long long)q_own, m);
  return qst_comm_alltoall(c, lists, recv, (size_t)q_own * (m + 1) * 8, stream);
}
