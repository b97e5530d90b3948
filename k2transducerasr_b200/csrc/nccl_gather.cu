// k2b_gather_results_nccl: the only collective of the path - one all-gather of the per-rank results (tokens, timestamps, counts,
// scores) over NVLink / NVSwitch, straight from the device buffers the search wrote, on the handle's stream.
//
// Streams are independent, so no collective runs inside the search; the gather exists for reporting.
// libnccl is NOT a link-time dependency of libk2b200.so (the C# host loads this library on machines without NCCL): it is
// dlopen'ed on first use - in a torch process that resolves to the libnccl.so.2 torch already mapped.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "k2b_internal.h"

namespace k2b {

namespace {
struct NcclUid { char internal[128]; };                   // ncclUniqueId
typedef struct ncclComm* nccl_comm_t;
typedef int (*fn_get_uid)(NcclUid*);
typedef int (*fn_init_rank)(nccl_comm_t*, int, NcclUid, int);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_void)(void);
typedef int (*fn_destroy)(nccl_comm_t);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_void group_start = nullptr, group_end = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
  std::string why;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib != nullptr) break;
    }
    if (a.lib == nullptr) { a.why = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "?"); return; }
    a.get_uid = reinterpret_cast<fn_get_uid>(dlsym(a.lib, "ncclGetUniqueId"));
    a.init_rank = reinterpret_cast<fn_init_rank>(dlsym(a.lib, "ncclCommInitRank"));
    a.all_gather = reinterpret_cast<fn_all_gather>(dlsym(a.lib, "ncclAllGather"));
    a.group_start = reinterpret_cast<fn_void>(dlsym(a.lib, "ncclGroupStart"));
    a.group_end = reinterpret_cast<fn_void>(dlsym(a.lib, "ncclGroupEnd"));
    a.destroy = reinterpret_cast<fn_destroy>(dlsym(a.lib, "ncclCommDestroy"));
    a.errstr = reinterpret_cast<fn_errstr>(dlsym(a.lib, "ncclGetErrorString"));
    if (!a.get_uid || !a.init_rank || !a.all_gather || !a.group_start || !a.group_end || !a.destroy) {
      a.why = "libnccl is missing one of the entry points this library binds";
      a.lib = nullptr;
    }
  });
  return a;
}

std::string nccl_msg(int rc) {
  NcclApi& a = api();
  return std::string("NCCL error ") + std::to_string(rc) + (a.errstr ? std::string(" (") + a.errstr(rc) + ")" : std::string());
}
constexpr int kNcclChar = 0;       // ncclInt8 / ncclChar: everything is gathered as bytes
}  // namespace

struct NcclState {
  nccl_comm_t comm = nullptr;
  int rank = 0, nranks = 1;
  // "async_gather": the gather runs on its own stream so that the next batch's search does not wait for it
  cudaStream_t side = nullptr;
  cudaEvent_t ev_src = nullptr, ev_done = nullptr;
  bool pending = false;                  // a gather on `side` that the handle's stream has not been ordered behind yet
};

void nccl_free(k2b_handle* h) {
  if (h->nccl == nullptr) return;
  if (h->nccl->side != nullptr) {
    cudaStreamSynchronize(h->nccl->side);
    cudaEventDestroy(h->nccl->ev_src);
    cudaEventDestroy(h->nccl->ev_done);
    cudaStreamDestroy(h->nccl->side);
  }
  if (h->nccl->comm != nullptr && api().destroy != nullptr) api().destroy(h->nccl->comm);
  delete h->nccl;
  h->nccl = nullptr;
}

// Orders the handle's stream behind an outstanding side-stream gather (no host synchronisation). Called by every entry point
// before it may overwrite result buffers a gather could still be reading; the cluster beam search defers it to its back-trace.
int32_t gather_join(k2b_handle* h) {
  if (h->nccl == nullptr || !h->nccl->pending) return K2B_OK;
  K2B_CUDA(h, cudaStreamWaitEvent(h->stream, h->nccl->ev_done, 0));
  h->nccl->pending = false;
  return K2B_OK;
}

}  // namespace k2b

using namespace k2b;

extern "C" {

int32_t k2b_nccl_unique_id(void* id128) {
  if (id128 == nullptr) return K2B_ERR_INVALID;
  NcclApi& a = api();
  if (a.lib == nullptr) return K2B_ERR_UNSUPPORTED;
  NcclUid u;
  const int rc = a.get_uid(&u);
  if (rc != 0) return K2B_ERR_CUDA;
  memcpy(id128, &u, sizeof(u));
  return K2B_OK;
}

int32_t k2b_nccl_init(k2b_handle* h, const void* id128, int32_t rank, int32_t nranks) {
  if (h == nullptr) return K2B_ERR_INVALID;
  if (h->poisoned) return K2B_ERR_STATE;
  K2B_CUDA(h, cudaSetDevice(h->cfg.device));
  if (id128 == nullptr || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, K2B_ERR_INVALID, "k2b_nccl_init: bad arguments");
  NcclApi& a = api();
  if (a.lib == nullptr) return fail(h, K2B_ERR_UNSUPPORTED, "k2b_nccl_init: " + a.why);
  nccl_free(h);
  h->nccl = new NcclState();
  NcclUid u;
  memcpy(&u, id128, sizeof(u));
  const int rc = a.init_rank(&h->nccl->comm, nranks, u, rank);
  if (rc != 0) { delete h->nccl; h->nccl = nullptr; return fail(h, K2B_ERR_CUDA, "k2b_nccl_init: " + nccl_msg(rc)); }
  h->nccl->rank = rank;
  h->nccl->nranks = nranks;
  return K2B_OK;
}

// Every rank passes the DEVICE buffers one fused search call filled for its own B streams ([B,cap] tokens / ts, [B] n / score;
// score may be NULL) and receives all ranks' results, rank-major, in DEVICE buffers of nranks times that size. Enqueued on the
// handle's stream (ordered behind the search that produced the inputs), or with "async_gather" on a side stream behind an event
// of the handle's stream; no host synchronisation.
int32_t k2b_gather_results_nccl(k2b_handle* h, const int64_t* tokens, const int32_t* ts, const int32_t* n, const float* score,
                                int32_t B, int32_t cap, int64_t* all_tokens, int32_t* all_ts, int32_t* all_n, float* all_score) {
  if (h == nullptr) return K2B_ERR_INVALID;
  if (h->poisoned) return K2B_ERR_STATE;
  K2B_CUDA(h, cudaSetDevice(h->cfg.device));
  if (h->nccl == nullptr) return fail(h, K2B_ERR_STATE, "k2b_gather_results_nccl: call k2b_nccl_init first");
  if (B < 0 || cap < 0 || (B > 0 && (!tokens || !ts || !n || !all_tokens || !all_ts || !all_n)) || ((score == nullptr) != (all_score == nullptr)))
    return fail(h, K2B_ERR_INVALID, "k2b_gather_results_nccl: bad arguments");
  if (B == 0) return K2B_OK;
  NcclApi& a = api();
  NcclState* st = h->nccl;
  nccl_comm_t c = st->comm;
  cudaStream_t on = h->stream;
  if (h->opt_async_gather != 0) {
    if (st->side == nullptr) {
      int lo = 0, hi = 0;
      K2B_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
      K2B_CUDA(h, cudaStreamCreateWithPriority(&st->side, cudaStreamNonBlocking, hi));
      K2B_CUDA(h, cudaEventCreateWithFlags(&st->ev_src, cudaEventDisableTiming));
      K2B_CUDA(h, cudaEventCreateWithFlags(&st->ev_done, cudaEventDisableTiming));
    }
    K2B_CUDA(h, cudaEventRecord(st->ev_src, h->stream));              // behind the search that produced the inputs
    K2B_CUDA(h, cudaStreamWaitEvent(st->side, st->ev_src, 0));
    on = st->side;
  } else {
    K2B_TRY(gather_join(h));
  }
  int rc = a.group_start();
  if (rc == 0 && cap > 0) rc = a.all_gather(tokens, all_tokens, sizeof(int64_t) * (size_t)B * cap, kNcclChar, c, on);
  if (rc == 0 && cap > 0) rc = a.all_gather(ts, all_ts, sizeof(int32_t) * (size_t)B * cap, kNcclChar, c, on);
  if (rc == 0) rc = a.all_gather(n, all_n, sizeof(int32_t) * (size_t)B, kNcclChar, c, on);
  if (rc == 0 && score != nullptr) rc = a.all_gather(score, all_score, sizeof(float) * (size_t)B, kNcclChar, c, on);
  const int rc2 = a.group_end();
  if (rc != 0 || rc2 != 0) return fail(h, K2B_ERR_CUDA, "k2b_gather_results_nccl: " + nccl_msg(rc != 0 ? rc : rc2));
  if (on == st->side) {
    K2B_CUDA(h, cudaEventRecord(st->ev_done, st->side));
    st->pending = true;
  }
  return K2B_OK;
}

int32_t k2b_gather_join(k2b_handle* h) {
  if (h == nullptr) return K2B_ERR_INVALID;
  if (h->poisoned) return K2B_ERR_STATE;
  K2B_CUDA(h, cudaSetDevice(h->cfg.device));
  return gather_join(h);
}

}  // extern "C"
