// The C ABI of libk2b200.so (include/k2b200.h): lifetime, weights, fine-grained proj calls and the host/device
// wrappers of the fused search loops. No CPU fallback lives here: without a CUDA device k2b_create fails.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "k2b_internal.h"

namespace k2b {

namespace {
thread_local std::string g_create_error;

__global__ void ctx_from_i64_kernel(const int64_t* __restrict__ y, int n, int blank, int32_t* __restrict__ ctx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (y != nullptr) {
    ctx[2 * i] = (int32_t)y[2 * i];
    ctx[2 * i + 1] = (int32_t)y[2 * i + 1];
  } else {  // ref OfflineProjOfTransducer.cs:97-110: null input -> n x {-1, blank}
    ctx[2 * i] = -1;
    ctx[2 * i + 1] = blank;
  }
}

// OnlineStream.Hyp arrives from the caller: ids outside the vocabulary must not become table addresses. Device-pointer callers get
// the ids clamped to blank and the error reported by the next k2b_sync / host-pointer call (dev_status[2]).
__global__ void check_hyp_kernel(int64_t* __restrict__ hyp, int B, int V, int blank, int allow_neg_tail, int* __restrict__ status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t c0 = hyp[2 * b], c1 = hyp[2 * b + 1];
  const bool bad0 = c0 < -1 || c0 >= V, bad1 = c1 >= V || c1 < (allow_neg_tail ? -1 : 0);
  if (bad0) hyp[2 * b] = blank;
  if (bad1) hyp[2 * b + 1] = blank;
  if (bad0 || bad1) atomicExch(status + 2, 1);
}
__global__ void fill_hyp_kernel(int64_t* __restrict__ hyp, int B, int c0, int c1) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) { hyp[2 * b] = c0; hyp[2 * b + 1] = c1; }
}
// Q6: frame of the batch's first emission = min over the streams of their first timestamp (T when nobody emits)
__global__ void first_emission_kernel(const int32_t* __restrict__ ts, const int32_t* __restrict__ n, int B, int cap, int T, int* __restrict__ tstar) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && n[b] > 0) atomicMin(tstar, ts[(size_t)b * cap]);
}
// streams that had not emitted when the batch first did are decoded again from frame tstar + 1 with the context {blank, blank}
// (pass 2); every other stream keeps its pass-1 result
__global__ void compat_select_kernel(int B, int cap, int tstar, int ts_off, const int64_t* __restrict__ tok2, const int32_t* __restrict__ ts2,
                                     const int32_t* __restrict__ n2, int64_t* __restrict__ tok, int32_t* __restrict__ ts, int32_t* __restrict__ n) {
  const int b = blockIdx.x;
  const bool redo = n[b] == 0 || ts[(size_t)b * cap] > tstar;
  __syncthreads();
  if (!redo) return;
  const int m = n2[b];
  for (int i = threadIdx.x; i < m && i < cap; i += blockDim.x) {
    tok[(size_t)b * cap + i] = tok2[(size_t)b * cap + i];
    ts[(size_t)b * cap + i] = ts2[(size_t)b * cap + i] + ts_off;
  }
  if (threadIdx.x == 0) n[b] = m;
}

void free_buf(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
}

void free_cluster_assets(k2b_handle* h) {
  if (h->wo_hi_img) cudaFree(h->wo_hi_img);
  if (h->wo_lo) cudaFree(h->wo_lo);
  if (h->wo_hi_rows) cudaFree(h->wo_hi_rows);
  if (h->bias_pad) cudaFree(h->bias_pad);
  if (h->dec_tab) cudaFree(h->dec_tab);
  h->wo_hi_img = nullptr; h->wo_lo = nullptr; h->wo_hi_rows = nullptr; h->bias_pad = nullptr; h->dec_tab = nullptr; h->dec_tab_state = 0;
  h->dec_tab_bytes = 0; h->dec_tab_build_ms = 0.f;
  h->tc_ready = false;
  if (h->we_hi_img) cudaFree(h->we_hi_img);
  if (h->we_lo_img) cudaFree(h->we_lo_img);
  h->we_hi_img = nullptr; h->we_lo_img = nullptr;
  h->enc_ready = false;
  if (h->wj_hi_img) cudaFree(h->wj_hi_img);
  if (h->wj_lo_img) cudaFree(h->wj_lo_img);
  h->wj_hi_img = nullptr; h->wj_lo_img = nullptr;
  h->wj_ready = false;
  if (h->wd_hi_img) cudaFree(h->wd_hi_img);
  if (h->wd_lo_img) cudaFree(h->wd_lo_img);
  h->wd_hi_img = nullptr; h->wd_lo_img = nullptr;
  h->wd_ready = false;
}

// join = false: the caller orders its stream behind an outstanding side-stream gather itself, in front of the first kernel that
// writes caller-owned result buffers (the gather may still be reading them)
int32_t enter(k2b_handle* h, bool join = true) {
  if (h == nullptr) return K2B_ERR_INVALID;
  if (h->poisoned) return K2B_ERR_STATE;
  cudaError_t e = cudaSetDevice(h->cfg.device);
  if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice", __FILE__, __LINE__);
  return join ? gather_join(h) : K2B_OK;
}

int32_t upload(k2b_handle* h, float** dst, const float* src, size_t n) {
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(dst), n * sizeof(float)));
  K2B_CUDA(h, cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  return K2B_OK;
}

int32_t need_weights(k2b_handle* h) {
  if (!h->weights_loaded) return fail(h, K2B_ERR_STATE, "weights not loaded (call k2b_load_weights first)");
  return K2B_OK;
}

// raw [n,E] -> projected [n,J] on device (the encoder_proj Linear)
int32_t encoder_proj_launch(k2b_handle* h, const float* raw, int n, float* out) {
  if (h->cfg.encoder_dim <= 0 || h->enc_w == nullptr) return fail(h, K2B_ERR_STATE, "encoder_proj weights not loaded (E == 0)");
  if (h->cfg.precision != K2B_PREC_FP32 && encproj_tc_supported(h)) return encoder_proj_tc(h, raw, n, out, false);
  GemmArgs a;
  a.M = n; a.N = h->cfg.joiner_dim; a.K = h->cfg.encoder_dim;
  a.A = raw; a.W = h->enc_w; a.bias = h->enc_b; a.C = out;
  return launch_gemm_simt(h, PRO_PLAIN, EPI_STORE, a);
}

// returns the device pointer of projected frames [B,T,J] for a fused call
int32_t frames_for_search(k2b_handle* h, const float* enc_dev, int enc_is_raw, int B, int T, const float** out) {
  if (!enc_is_raw) { *out = enc_dev; return K2B_OK; }
  const size_t n = (size_t)B * T;
  K2B_TRY(ensure(h, h->ws_encproj, sizeof(float) * n * h->cfg.joiner_dim));
  K2B_TRY(encoder_proj_launch(h, enc_dev, (int)n, static_cast<float*>(h->ws_encproj.p)));
  *out = static_cast<const float*>(h->ws_encproj.p);
  return K2B_OK;
}

// Device-pointer search on raw frames with the encoder_proj GEMM hidden under the search, ONE cluster-kernel launch: the frames are
// projected in time chunks of 32 on a low-priority side stream (on the SMs the cluster kernel leaves free); after each chunk a
// one-thread kernel releases "chunk c is there" (an epoch in la_flags[c]); the running search polls the flag of a chunk in front of
// that chunk's first frame load. Only the first chunk's GEMM is waited for with an event in front of the launch (it has the whole
// GPU). The cluster kernel must leave SMs free while it polls, so this form is used only when its grid is at least 16 SMs smaller
// than the device (cfg2: 128 of 148). Two projected-frame buffers are taken in turns; the side stream waits for the last search
// that read the buffer it is about to fill.
// With k2b_set_option("inputs_complete", 1) the caller promises that its frames are complete in memory at call time (an encoder
// that ran on another stream and was synchronised): the side stream is then not ordered behind the handle's stream, so when calls
// follow each other the next call's chunks are projected under the current search and its launch finds them ready.
constexpr int kLaFlags = 1024;
static int32_t lookahead_init(k2b_handle* h) {
  int lo = 0, hi = 0;
  K2B_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
  K2B_CUDA(h, cudaStreamCreateWithPriority(&h->la_stream, cudaStreamNonBlocking, lo));
  K2B_CUDA(h, cudaEventCreateWithFlags(&h->la_ev_a, cudaEventDisableTiming));
  K2B_CUDA(h, cudaEventCreateWithFlags(&h->la_ev_b, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) K2B_CUDA(h, cudaEventCreateWithFlags(&h->la_ev_done[i], cudaEventDisableTiming));
  K2B_CUDA(h, cudaEventCreateWithFlags(&h->la_ev_mark, cudaEventDisableTiming));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->la_flags), sizeof(int) * kLaFlags));
  K2B_CUDA(h, cudaMemset(h->la_flags, 0, sizeof(int) * kLaFlags));
  return K2B_OK;
}
// Frames per time chunk: the chunk's GEMM runs on the F SMs the search leaves free, in waves of F tiles of 128 rows x 256 columns,
// so a chunk should be a whole number of waves (cfg2: 25 frames = 100 tiles = 5 waves of 20; 32 frames would be 6.4 -> 7 waves)
static int chunk_frames(const k2b_handle* h, int B, int T, int K) {
  if (h->opt_la_frames > 0) return h->opt_la_frames;
  const int F = h->sm_count - cluster_grid_ctas(h, B, K), ntn = (h->cfg.joiner_dim + 255) / 256;
  int best = 32;
  double best_cost = 1e30;
  for (int L = 16; L <= 48; ++L) {
    const int tiles = (int)(((long long)B * L + 127) / 128) * ntn, waves = (tiles + F - 1) / F;
    const int rem = T % L, rem_tiles = (int)(((long long)B * rem + 127) / 128) * ntn;
    const int nfull = T / L;
    const double cost = ((double)nfull * (waves + 0.3) + (rem ? (rem_tiles + F - 1) / F + 0.3 : 0.0)) / T;   // waves per frame
    if (cost < best_cost - 1e-9) { best_cost = cost; best = L; }
  }
  return best;
}
static bool flagged_search_possible(const k2b_handle* h, int B, int T, int K) {
  if (cluster_grid_ctas(h, B, K) + 16 > h->sm_count) return false;
  const int L = chunk_frames(h, B, T, K);
  return (T + L - 1) / L <= kLaFlags;
}
static int32_t beam_cluster_flagged(k2b_handle* h, const float* enc, int B, int T, int K, int64_t* tokens, int32_t* ts, int32_t* n_out,
                                    float* score, int cap, int extra_mask) {
  const size_t J = h->cfg.joiner_dim;
  if (h->la_stream == nullptr) K2B_TRY(lookahead_init(h));
  const int X = (int)(h->la_turn++ & 1u);
  DevBuf& buf = h->ws_encproj_la[X];
  const size_t need = sizeof(float) * (size_t)B * T * J;
  if (need > buf.bytes) {
    K2B_CUDA(h, cudaStreamSynchronize(h->la_stream));
    K2B_TRY(ensure(h, buf, need));
    h->la_done_valid[X] = false;
  }
  const bool cold = !h->enc_ready;
  K2B_TRY(ensure_encproj_assets(h));        // (packs the weight images on the compute stream the first time)
  if (cold) K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  const size_t NK = (size_t)B * K;
  const size_t a4 = (NK * 4 + 255) & ~size_t(255), ab = ((size_t)B * 4 + 255) & ~size_t(255);
  K2B_TRY(ensure(h, h->ws_state, 2 * a4 + 2 * ab));
  K2B_TRY(ensure(h, h->ws_bp, sizeof(int32_t) * (size_t)B * T * K));
  char* q = static_cast<char*>(h->ws_state.p);
  float* fin_lp = reinterpret_cast<float*>(q); q += a4;
  int32_t* fin_len = reinterpret_cast<int32_t*>(q); q += a4;
  int32_t* fin_nlive = reinterpret_cast<int32_t*>(q); q += ab;
  const bool need_lp = score != nullptr;
  if (score == nullptr) score = reinterpret_cast<float*>(q);
  int32_t* bp = static_cast<int32_t*>(h->ws_bp.p);
  float* encE = static_cast<float*>(buf.p);
  const int epoch = ++h->la_epoch;
  if (h->opt_inputs_complete == 0) {        // the frames (and nothing else of this call) are ordered on the handle's stream
    K2B_CUDA(h, cudaEventRecord(h->la_ev_b, h->stream));
    K2B_CUDA(h, cudaStreamWaitEvent(h->la_stream, h->la_ev_b, 0));
  } else if (h->la_mark_valid) {            // not before the previous call's search has been launched (see la_ev_mark)
    K2B_CUDA(h, cudaStreamWaitEvent(h->la_stream, h->la_ev_mark, 0));
  }
  if (h->la_done_valid[X]) K2B_CUDA(h, cudaStreamWaitEvent(h->la_stream, h->la_ev_done[X], 0));
  const int L = chunk_frames(h, B, T, K);
  for (int c = 0, t0 = 0; t0 < T; t0 += L, ++c) {
    const int tc = T - t0 < L ? T - t0 : L;
    cudaStream_t keep = h->stream;
    h->stream = h->la_stream;
    int32_t st = encoder_proj_tc(h, enc, B * tc, encE, true, tc, T, t0, T);
    if (st == K2B_OK) st = cluster_set_ready(h, h->la_flags + c, epoch);
    h->stream = keep;
    K2B_TRY(st);
    if (c == 0) K2B_CUDA(h, cudaEventRecord(h->la_ev_a, h->la_stream));
  }
  // the search is enqueued behind ALL the projections in host order (on the device it only waits for the first chunk): tools
  // that run one kernel at a time in launch order (ncu, CUDA_LAUNCH_BLOCKING) then find every flag set instead of a search that
  // polls for a kernel which cannot start
  K2B_CUDA(h, cudaStreamWaitEvent(h->stream, h->la_ev_a, 0));
  K2B_CUDA(h, cudaEventRecord(h->la_ev_mark, h->stream));
  h->la_mark_valid = true;
  K2B_TRY(beam_cluster_dev(h, encE, B, T, K, bp, fin_lp, fin_len, fin_nlive, extra_mask, nullptr, nullptr, 0, 0, 0, nullptr, nullptr,
                           need_lp, h->la_flags, epoch, L));
  K2B_CUDA(h, cudaEventRecord(h->la_ev_done[X], h->stream));
  h->la_done_valid[X] = true;
  K2B_TRY(gather_join(h));                 // the back-trace is the first kernel that writes the caller's result buffers
  return beam_backtrace_dev(h, B, K, T, fin_lp, fin_len, fin_nlive, bp, tokens, ts, n_out, score, cap, K > 1 ? cluster_cst(h, B, K) : nullptr);
}

// modified_beam_search on the persistent cluster kernel (tcgen05 precisions): frames -> exp(2x) (fused into the
// encoder_proj epilogue when the frames are raw), one launch for the whole time loop, then the back-trace.
int32_t beam_cluster_path(k2b_handle* h, const float* enc, int enc_is_raw, int B, int T, int K, int64_t* tokens, int32_t* ts,
                          int32_t* n_out, float* score, int cap, int extra_mask = -1, int64_t* hyp_inout = nullptr) {
  K2B_TRY(ensure_cluster_assets(h));
  const size_t n = (size_t)B * T, J = h->cfg.joiner_dim;
  K2B_TRY(ensure(h, h->ws_encproj, sizeof(float) * n * J));
  float* encE = static_cast<float*>(h->ws_encproj.p);
  // Raw frames resident on the device, a batch that fills the cluster kernel's 128 SMs: the encoder_proj GEMM of the SECOND half
  // of the frames runs on a side stream (on the SMs the search leaves free) under the search of the first half; the search is two
  // launches with the hypothesis state carried through global memory, as in the host-pointer call. cfg2: 1.43 -> ~1.34 ms per batch.
  const bool two = enc_is_raw && hyp_inout == nullptr && encproj_tc_supported(h) && h->cfg.encoder_dim > 0 && h->enc_w != nullptr &&
                   (h->opt_dev_chunks < 0 ? (T >= 64 && (long long)B * K >= 512) : h->opt_dev_chunks >= 2) && !h->profile_on;
  // measured on cfg2 (ms per batch, end of round 2): one launch behind the whole projection 1.395, two launches with events 1.365,
  // one polling launch 1.314, and 1.253 when its projections may run ahead of the handle's stream ("inputs_complete"); chunks of
  // 16 / 64 frames instead of 32: 1.67 / 1.35 (1.54 / 1.29 with "inputs_complete")
  if (two && (h->opt_dev_chunks < 0 || h->opt_dev_chunks == 3) && flagged_search_possible(h, B, T, K))
    return beam_cluster_flagged(h, enc, B, T, K, tokens, ts, n_out, score, cap, extra_mask);
  if (two) {
    const int tA = T / 2, tB = T - tA;
    if (h->copy_stream == nullptr) {
      // lowest priority: when the search's CTAs become ready, the SMs that projection tiles free go to them first
      int lo = 0, hi = 0;
      K2B_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
      K2B_CUDA(h, cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, lo));
      for (int i = 0; i < 2; ++i) {
        K2B_CUDA(h, cudaEventCreateWithFlags(&h->ev_ready[i], cudaEventDisableTiming));
        K2B_CUDA(h, cudaEventCreateWithFlags(&h->ev_free[i], cudaEventDisableTiming));
      }
    }
    const size_t NK2 = (size_t)B * K;
    const size_t a4 = (NK2 * 4 + 255) & ~size_t(255), a8 = (NK2 * 8 + 255) & ~size_t(255), ab = ((size_t)B * 4 + 255) & ~size_t(255);
    K2B_TRY(ensure(h, h->ws_state, 2 * a4 + ab + a8 + a8 + ab));
    K2B_TRY(ensure(h, h->ws_bp, sizeof(int32_t) * (size_t)B * T * K));
    K2B_TRY(ensure_encproj_assets(h));      // (packs the weight images on the compute stream the first time)
    char* q = static_cast<char*>(h->ws_state.p);
    float* fin_lp = reinterpret_cast<float*>(q); q += a4;
    int32_t* fin_len = reinterpret_cast<int32_t*>(q); q += a4;
    int32_t* fin_nlive = reinterpret_cast<int32_t*>(q); q += ab;
    int32_t* io_ctx = reinterpret_cast<int32_t*>(q); q += a8;
    unsigned long long* io_hash = reinterpret_cast<unsigned long long*>(q); q += a8;
    const bool need_lp2 = score != nullptr;
    if (score == nullptr) score = reinterpret_cast<float*>(q);
    int32_t* bp = static_cast<int32_t*>(h->ws_bp.p);
    // the side stream starts behind the first half's projection (the caller's frames are ready, the workspaces free, and the two
    // GEMMs do not compete)
    K2B_TRY(encoder_proj_tc(h, enc, B * tA, encE, true, tA, T, 0, T));
    K2B_CUDA(h, cudaEventRecord(h->ev_free[0], h->stream));
    K2B_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_free[0], 0));
    K2B_TRY(beam_cluster_dev(h, encE, B, tA, K, bp, fin_lp, fin_len, fin_nlive, extra_mask, nullptr, nullptr, 0, T, 0, io_ctx, io_hash, need_lp2));
    {
      cudaStream_t keep = h->stream;
      h->stream = h->copy_stream;
      const int32_t st = encoder_proj_tc(h, enc, B * tB, encE, true, tB, T, tA, T);
      h->stream = keep;
      K2B_TRY(st);
    }
    K2B_CUDA(h, cudaEventRecord(h->ev_ready[0], h->copy_stream));
    K2B_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_ready[0], 0));
    K2B_TRY(beam_cluster_dev(h, encE, B, tB, K, bp, fin_lp, fin_len, fin_nlive, extra_mask, nullptr, nullptr, tA, T, 1, io_ctx, io_hash, need_lp2));
    K2B_TRY(gather_join(h));               // the back-trace is the first kernel that writes the caller's result buffers
    return beam_backtrace_dev(h, B, K, T, fin_lp, fin_len, fin_nlive, bp, tokens, ts, n_out, score, cap, K > 1 ? cluster_cst(h, B, K) : nullptr);
  }
  if (enc_is_raw) {
    if (h->cfg.encoder_dim <= 0 || h->enc_w == nullptr) return fail(h, K2B_ERR_STATE, "encoder_proj weights not loaded (E == 0)");
    if (encproj_tc_supported(h)) {
      K2B_TRY(encoder_proj_tc(h, enc, (int)n, encE, true));
    } else {
      GemmArgs g;
      g.M = (int)n; g.N = (int)J; g.K = h->cfg.encoder_dim;
      g.A = enc; g.W = h->enc_w; g.bias = h->enc_b; g.C = encE;
      K2B_TRY(launch_gemm_simt(h, PRO_PLAIN, EPI_EXP2X, g));
    }
  } else {
    K2B_TRY(exp2x_frames(h, enc, encE, n * J));
  }
  const size_t NK = (size_t)B * K;
  const size_t bytes = ((NK * 4 + 255) & ~size_t(255)) * 2 + (((size_t)B * 4 + 255) & ~size_t(255)) * 2;
  K2B_TRY(ensure(h, h->ws_state, bytes));
  K2B_TRY(ensure(h, h->ws_bp, sizeof(int32_t) * (size_t)B * T * K));
  char* p = static_cast<char*>(h->ws_state.p);
  float* fin_lp = reinterpret_cast<float*>(p); p += (NK * 4 + 255) & ~size_t(255);
  int32_t* fin_len = reinterpret_cast<int32_t*>(p); p += (NK * 4 + 255) & ~size_t(255);
  int32_t* fin_nlive = reinterpret_cast<int32_t*>(p); p += ((size_t)B * 4 + 255) & ~size_t(255);
  const bool need_lp = score != nullptr;                            // greedy callers have no use for the score
  if (score == nullptr) score = reinterpret_cast<float*>(p);
  int32_t* bp = static_cast<int32_t*>(h->ws_bp.p);
  K2B_TRY(beam_cluster_dev(h, encE, B, T, K, bp, fin_lp, fin_len, fin_nlive, extra_mask, hyp_inout, hyp_inout, 0, 0, 0, nullptr, nullptr,
                           need_lp));
  K2B_TRY(gather_join(h));
  return beam_backtrace_dev(h, B, K, T, fin_lp, fin_len, fin_nlive, bp, tokens, ts, n_out, score, cap, K > 1 ? cluster_cst(h, B, K) : nullptr);
}

// Time chunks of a host-pointer call: short chunks at both ends, equal long ones between them. The first chunk's copy and the last
// chunk's search are the two pieces nothing can hide (copy-bound: the tail; search-bound: the head), so both are kept at 8 frames;
// the chunks in between are long so that the ~25 us of set-up per launch stays small against them.
static std::vector<int> chunk_schedule(int T, int nchunk_fixed, int big) {
  std::vector<int> v;
  if (nchunk_fixed >= 1) {
    const int Tc = (T + nchunk_fixed - 1) / nchunk_fixed;
    for (int t0 = 0; t0 < T; t0 += Tc) v.push_back((T - t0) < Tc ? (T - t0) : Tc);
    return v;
  }
  std::vector<int> tail;
  int rest = T;
  if (T >= 32) { v.push_back(8); tail.push_back(8); rest -= 16; }
  if (T >= 128) { v.push_back(16); tail.insert(tail.begin(), 16); rest -= 32; }
  const int nmid = (rest + big - 1) / big;
  for (int i = 0; i < nmid; ++i) v.push_back(rest / nmid + (i < rest % nmid ? 1 : 0));
  v.insert(v.end(), tail.begin(), tail.end());
  return v;
}

// Host-pointer modified_beam_search with the input copy hidden behind the search: the batch is cut into time chunks;
// chunk c+1 crosses PCIe (strided 2-D copy into a compact [B,Tc,E] staging buffer, on a copy stream) while chunk c is
// projected and decoded (one cluster-kernel launch per chunk, hypothesis state carried through global memory).
int32_t beam_cluster_pipelined(k2b_handle* h, const float* enc_host, int enc_is_raw, int B, int T, int K, int64_t* tokens, int32_t* ts,
                               int32_t* n_out, float* score, int cap) {
  K2B_TRY(ensure_cluster_assets(h));
  const int J = h->cfg.joiner_dim;
  const int E = enc_is_raw ? h->cfg.encoder_dim : J;      // width of what crosses PCIe: raw frames, or the seam's projected ones
  // measured on cfg2 (PCIe-bound, 196 MB in) with equal chunks: 4 chunks 4.04 ms, 8 chunks 3.85 ms, 12 chunks 3.81 ms per batch - the
  // un-overlapped tail (last chunk's projection + search) shrinks, each extra launch costs ~30 us; hence the schedule above
  const std::vector<int> sched = chunk_schedule(T, h->opt_pipe_chunks, 32);
  int Tc = 1;
  for (int x : sched) Tc = x > Tc ? x : Tc;
  if (h->copy_stream == nullptr) {
    K2B_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      K2B_CUDA(h, cudaEventCreateWithFlags(&h->ev_ready[i], cudaEventDisableTiming));
      K2B_CUDA(h, cudaEventCreateWithFlags(&h->ev_free[i], cudaEventDisableTiming));
    }
  }
  const size_t buf_bytes = ((sizeof(float) * (size_t)B * Tc * E) + 255) & ~size_t(255);
  K2B_TRY(ensure(h, h->ws_in, 2 * buf_bytes));
  K2B_TRY(ensure(h, h->ws_encproj, sizeof(float) * (size_t)B * T * J));
  const size_t NK = (size_t)B * K;
  const size_t a4 = (NK * 4 + 255) & ~size_t(255), a8 = (NK * 8 + 255) & ~size_t(255), ab = ((size_t)B * 4 + 255) & ~size_t(255);
  K2B_TRY(ensure(h, h->ws_state, 2 * a4 + ab + a8 + a8));
  K2B_TRY(ensure(h, h->ws_bp, sizeof(int32_t) * (size_t)B * T * K));
  char* p = static_cast<char*>(h->ws_state.p);
  float* fin_lp = reinterpret_cast<float*>(p); p += a4;
  int32_t* fin_len = reinterpret_cast<int32_t*>(p); p += a4;
  int32_t* fin_nlive = reinterpret_cast<int32_t*>(p); p += ab;
  int32_t* io_ctx = reinterpret_cast<int32_t*>(p); p += a8;
  unsigned long long* io_hash = reinterpret_cast<unsigned long long*>(p);
  float* encE = static_cast<float*>(h->ws_encproj.p);
  int32_t* bp = static_cast<int32_t*>(h->ws_bp.p);
  // the staging buffers may still be in use by earlier work on the compute stream
  K2B_CUDA(h, cudaEventRecord(h->ev_free[0], h->stream));
  K2B_CUDA(h, cudaEventRecord(h->ev_free[1], h->stream));
  int c = 0, t0 = 0;
  for (size_t ci = 0; ci < sched.size(); t0 += sched[ci], ++ci, ++c) {
    const int tc = sched[ci], sb = c & 1;
    float* stage = reinterpret_cast<float*>(static_cast<char*>(h->ws_in.p) + (size_t)sb * buf_bytes);
    K2B_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_free[sb], 0));
    K2B_TRY(h2d_rows(h, stage, sizeof(float) * (size_t)tc * E, enc_host + (size_t)t0 * E, sizeof(float) * (size_t)T * E,
                     sizeof(float) * (size_t)tc * E, (size_t)B, h->copy_stream));
    K2B_CUDA(h, cudaEventRecord(h->ev_ready[sb], h->copy_stream));
    K2B_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_ready[sb], 0));
    if (enc_is_raw) K2B_TRY(encoder_proj_tc(h, stage, B * tc, encE, true, tc, T, t0));
    else K2B_TRY(exp2x_frames_chunk(h, stage, encE, B, tc, T, t0));
    K2B_CUDA(h, cudaEventRecord(h->ev_free[sb], h->stream));
    K2B_TRY(beam_cluster_dev(h, encE, B, tc, K, bp, fin_lp, fin_len, fin_nlive, -1, nullptr, nullptr, t0, T, c > 0 ? 1 : 0, io_ctx, io_hash));
  }
  return beam_backtrace_dev(h, B, K, T, fin_lp, fin_len, fin_nlive, bp, tokens, ts, n_out, score, cap, K > 1 ? cluster_cst(h, B, K) : nullptr);
}

// projected frames of time chunk [t0, t0 + tc) of every stream, written into a [B,T,J] array (tcgen05 GEMM, or the CUDA-core one
// into a compact buffer followed by a strided copy when J is not a multiple of the tensor-core tile)
static int32_t encoder_proj_chunk(k2b_handle* h, const float* stage, int B, int tc, float* encP, int T, int t0) {
  const int E = h->cfg.encoder_dim, J = h->cfg.joiner_dim;
  if (encproj_tc_supported(h)) return encoder_proj_tc(h, stage, B * tc, encP, false, tc, T, t0);
  K2B_TRY(ensure(h, h->ws_x, sizeof(float) * (size_t)B * tc * J));
  GemmArgs g;
  g.M = B * tc; g.N = J; g.K = E;
  g.A = stage; g.W = h->enc_w; g.bias = h->enc_b; g.C = static_cast<float*>(h->ws_x.p);
  K2B_TRY(launch_gemm_simt(h, PRO_PLAIN, EPI_STORE, g));
  K2B_CUDA(h, cudaMemcpy2DAsync(encP + (size_t)t0 * J, sizeof(float) * (size_t)T * J, h->ws_x.p, sizeof(float) * (size_t)tc * J,
                                sizeof(float) * (size_t)tc * J, (size_t)B, cudaMemcpyDeviceToDevice, h->stream));
  return K2B_OK;
}

// The same for the engines of beam_dev that can be stepped in time chunks (persistent beam kernel: large vocabularies): chunk c+1
// crosses PCIe while chunk c is projected and searched; the hypothesis state stays in the workspaces between the chunks.
int32_t beam_chunked_host(k2b_handle* h, const float* enc_host, int enc_is_raw, int B, int T, int K, int64_t* tokens, int32_t* ts,
                          int32_t* n_out, float* score, int cap) {
  const int J = h->cfg.joiner_dim, E = enc_is_raw ? h->cfg.encoder_dim : J;
  // each chunk is one persistent launch (~20 us of set-up); measured on cfg4 (131 MB in): 3 chunks 6.06 ms, 5 chunks 5.82 ms,
  // 8 chunks 5.64 ms, 12 chunks 5.67 ms per batch
  const std::vector<int> sched = chunk_schedule(T, h->opt_pipe_chunks, 32);
  int Tc = 1;
  for (int x : sched) Tc = x > Tc ? x : Tc;
  if (h->copy_stream == nullptr) {
    K2B_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      K2B_CUDA(h, cudaEventCreateWithFlags(&h->ev_ready[i], cudaEventDisableTiming));
      K2B_CUDA(h, cudaEventCreateWithFlags(&h->ev_free[i], cudaEventDisableTiming));
    }
  }
  const size_t buf_bytes = ((sizeof(float) * (size_t)B * Tc * E) + 255) & ~size_t(255);
  if (enc_is_raw) K2B_TRY(ensure(h, h->ws_in, 2 * buf_bytes));
  K2B_TRY(ensure(h, h->ws_encproj, sizeof(float) * (size_t)B * T * J));
  float* encP = static_cast<float*>(h->ws_encproj.p);
  K2B_CUDA(h, cudaEventRecord(h->ev_free[0], h->stream));
  K2B_CUDA(h, cudaEventRecord(h->ev_free[1], h->stream));
  int c = 0, t0 = 0;
  for (size_t ci = 0; ci < sched.size(); t0 += sched[ci], ++ci, ++c) {
    const int tc = sched[ci], sb = c & 1;
    K2B_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_free[sb], 0));
    if (enc_is_raw) {
      float* stage = reinterpret_cast<float*>(static_cast<char*>(h->ws_in.p) + (size_t)sb * buf_bytes);
      K2B_TRY(h2d_rows(h, stage, sizeof(float) * (size_t)tc * E, enc_host + (size_t)t0 * E, sizeof(float) * (size_t)T * E,
                       sizeof(float) * (size_t)tc * E, (size_t)B, h->copy_stream));
      K2B_CUDA(h, cudaEventRecord(h->ev_ready[sb], h->copy_stream));
      K2B_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_ready[sb], 0));
      K2B_TRY(encoder_proj_chunk(h, stage, B, tc, encP, T, t0));
    } else {
      // the seam's own payload, already projected: the chunk's columns go straight to their place in [B,T,J] (no staging buffer; the
      // array as a whole is free once the work recorded in ev_free has finished)
      K2B_TRY(h2d_rows(h, encP + (size_t)t0 * J, sizeof(float) * (size_t)T * J, enc_host + (size_t)t0 * J, sizeof(float) * (size_t)T * J,
                       sizeof(float) * (size_t)tc * J, (size_t)B, h->copy_stream));
      K2B_CUDA(h, cudaEventRecord(h->ev_ready[sb], h->copy_stream));
      K2B_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_ready[sb], 0));
    }
    K2B_CUDA(h, cudaEventRecord(h->ev_free[sb], h->stream));
    K2B_TRY(beam_dev(h, encP + (size_t)t0 * J, B, tc, K, tokens, ts, n_out, score, cap, -1, nullptr, false, t0, T));
  }
  return K2B_OK;
}

struct OutStage {
  int64_t* tokens;
  int32_t* ts;
  int32_t* n;
  float* score;
  int64_t* hyp;
  int32_t* aux_a;  // frame_offset / trailing
  int32_t* aux_b;
  int64_t* prev;
};

int32_t stage_out(k2b_handle* h, int B, int cap, OutStage* o) {
  const size_t tok_b = (sizeof(int64_t) * (size_t)B * cap + 255) & ~size_t(255);
  const size_t ts_b = (sizeof(int32_t) * (size_t)B * cap + 255) & ~size_t(255);
  const size_t per_b = (sizeof(int64_t) * (size_t)B * 2 + 255) & ~size_t(255);
  K2B_TRY(ensure(h, h->ws_out, tok_b + ts_b + 6 * per_b));
  char* p = static_cast<char*>(h->ws_out.p);
  o->tokens = reinterpret_cast<int64_t*>(p); p += tok_b;
  o->ts = reinterpret_cast<int32_t*>(p); p += ts_b;
  o->n = reinterpret_cast<int32_t*>(p); p += per_b;
  o->score = reinterpret_cast<float*>(p); p += per_b;
  o->hyp = reinterpret_cast<int64_t*>(p); p += per_b;
  o->aux_a = reinterpret_cast<int32_t*>(p); p += per_b;
  o->aux_b = reinterpret_cast<int32_t*>(p); p += per_b;
  o->prev = reinterpret_cast<int64_t*>(p); p += per_b;
  return K2B_OK;
}

int32_t check_search_args(k2b_handle* h, const void* enc, int B, int T, int cap, const void* tokens, const void* ts,
                          const void* n_out, const char* who) {
  if (B < 0 || T < 0) return fail(h, K2B_ERR_INVALID, std::string(who) + ": negative B or T");
  if (B > 0 && T > 0 && enc == nullptr) return fail(h, K2B_ERR_INVALID, std::string(who) + ": enc is NULL");
  if (B > 0 && (tokens == nullptr || ts == nullptr || n_out == nullptr))
    return fail(h, K2B_ERR_INVALID, std::string(who) + ": output buffer is NULL");
  if (cap < T) return fail(h, K2B_ERR_INVALID, std::string(who) + ": cap must be >= T (one symbol per frame at most)");
  return K2B_OK;
}

}  // namespace

int32_t fail(k2b_handle* h, int32_t code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

int32_t cuda_fail(k2b_handle* h, cudaError_t e, const char* what, const char* file, int line) {
  std::string m = std::string("CUDA error ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ") at " + what +
                  " [" + file + ":" + std::to_string(line) + "]";
  if (h) { h->err = m; h->poisoned = true; } else { g_create_error = m; }
  return K2B_ERR_CUDA;
}

int32_t ensure(k2b_handle* h, DevBuf& b, size_t bytes) {
  if (bytes <= b.bytes) return K2B_OK;
  const size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
  K2B_CUDA(h, cudaMalloc(&b.p, want));
  b.bytes = want;
  return K2B_OK;
}

}  // namespace k2b

using namespace k2b;

extern "C" {

int32_t k2b_abi_version(void) { return K2B_ABI_VERSION; }

const char* k2b_last_error(const k2b_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int32_t k2b_create(const k2b_config* cfg, k2b_handle** out) {
  if (out == nullptr) return fail(nullptr, K2B_ERR_INVALID, "k2b_create: out is NULL");
  *out = nullptr;
  if (cfg == nullptr) return fail(nullptr, K2B_ERR_INVALID, "k2b_create: cfg is NULL");
  if (cfg->struct_size != (int32_t)sizeof(k2b_config))
    return fail(nullptr, K2B_ERR_INVALID, "k2b_create: struct_size does not match this library's k2b_config");
  if (cfg->context_size != 2)
    return fail(nullptr, K2B_ERR_UNSUPPORTED, "k2b_create: context_size must be 2 (the reference only works with 2)");
  if (cfg->vocab_size < 1 || cfg->joiner_dim < 16 || cfg->decoder_dim < 16 || cfg->joiner_dim % 16 || cfg->decoder_dim % 16 ||
      cfg->encoder_dim < 0 || cfg->encoder_dim % 16)
    return fail(nullptr, K2B_ERR_INVALID, "k2b_create: V >= 1 and J, D, E multiples of 16 are required");
  if (cfg->vocab_size >= (1 << 27)) return fail(nullptr, K2B_ERR_INVALID, "k2b_create: vocab_size too large");
  if (cfg->max_beam > kMaxBeam) return fail(nullptr, K2B_ERR_INVALID, "k2b_create: max_beam > 8");
  if (cfg->precision < K2B_PREC_FP32 || cfg->precision > K2B_PREC_BF16)
    return fail(nullptr, K2B_ERR_INVALID, "k2b_create: unknown precision");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail(nullptr, K2B_ERR_CUDA, std::string("k2b_create: no CUDA device (") + cudaGetErrorString(e) +
                                          "); this library has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, K2B_ERR_INVALID, "k2b_create: device ordinal out of range");
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice", __FILE__, __LINE__);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, cfg->device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties", __FILE__, __LINE__);
  if (prop.major != 10)
    return fail(nullptr, K2B_ERR_UNSUPPORTED, "k2b_create: this build targets sm_100a (Blackwell B200) only");
  k2b_handle* h = new k2b_handle();
  h->cfg = *cfg;
  h->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaStreamCreate", __FILE__, __LINE__); }
  h->stream = h->own_stream;
  e = cudaMalloc(reinterpret_cast<void**>(&h->dev_status), 4 * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(h->dev_status, 0, 4 * sizeof(int));
  if (e != cudaSuccess) { cudaStreamDestroy(h->own_stream); delete h; return cuda_fail(nullptr, e, "cudaMalloc(status)", __FILE__, __LINE__); }
  // engine switches: the environment only seeds them (tools set it before creating the handle); k2b_set_option changes them later
  auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e != nullptr ? atoi(e) : dflt; };
  h->opt_pipe_chunks = env_int("K2B_PIPE_CHUNKS", 0);
  h->opt_no_mega = env_int("K2B_NO_MEGA", 0);
  h->opt_unfused_step = env_int("K2B_UNFUSED_STEP", 0);
  h->opt_greedy_persistent = env_int("K2B_GREEDY_PERSISTENT", -1);
  h->opt_pair = env_int("K2B_PAIR", 0);
  h->prof_which = env_int("K2B_PROF_WHICH", 0);
  *out = h;
  return K2B_OK;
}

int32_t k2b_set_option(k2b_handle* h, const char* name, int32_t value) {
  K2B_TRY(enter(h));
  if (name == nullptr) return fail(h, K2B_ERR_INVALID, "k2b_set_option: name is NULL");
  const std::string n(name);
  if (n == "pipe_chunks") { if (value < 0 || value > 64) return fail(h, K2B_ERR_INVALID, "pipe_chunks: 0 (auto) .. 64"); h->opt_pipe_chunks = value; }
  else if (n == "no_mega") h->opt_no_mega = value;
  else if (n == "unfused_step") h->opt_unfused_step = value;
  else if (n == "greedy_persistent") h->opt_greedy_persistent = value;
  else if (n == "pair") h->opt_pair = value;
  else if (n == "wh_tmem_kb") h->opt_wh_tmem = value;
  else if (n == "dev_chunks") h->opt_dev_chunks = value;
  else if (n == "single_greedy") h->opt_single_greedy = value;
  else if (n == "tagged_records") h->opt_tagged_records = value;
  else if (n == "ctc_one_kernel") h->opt_ctc_one_kernel = value;
  else if (n == "prof_which") h->prof_which = value;
  else if (n == "async_d2h") h->opt_async_d2h = value;
  else if (n == "async_gather") h->opt_async_gather = value;
  else if (n == "inputs_complete") h->opt_inputs_complete = value;
  else if (n == "dev_chunk_frames") { if (value < 0 || value > 4096 || value == 1) return fail(h, K2B_ERR_INVALID, "dev_chunk_frames: 0 (auto), 2 .. 4096"); h->opt_la_frames = value; }
  else if (n == "copy_threads") {
    if (value < -1 || value > 64) return fail(h, K2B_ERR_INVALID, "copy_threads: -1 (auto) .. 64");
    if (value != h->opt_copy_threads) {                        // the pool is rebuilt with the new size by the next pageable copy
      K2B_CUDA(h, cudaStreamSynchronize(h->stream));
      if (h->copy_stream != nullptr) K2B_CUDA(h, cudaStreamSynchronize(h->copy_stream));
      host_stage_free(h);
    }
    h->opt_copy_threads = value;
  }
  else if (n == "max_sym_per_frame") {
    if (value < 1 || value > 16) return fail(h, K2B_ERR_INVALID, "max_sym_per_frame: 1 .. 16");
    h->max_sym_per_frame = value;
  } else if (n == "cluster_timing") {
    if (value == 0 && h->cluster_timing != nullptr) {
      K2B_CUDA(h, cudaStreamSynchronize(h->stream));
      cudaFree(h->cluster_timing);
      h->cluster_timing = nullptr;
    }
  } else return fail(h, K2B_ERR_INVALID, "k2b_set_option: unknown option " + n);
  return K2B_OK;
}

int32_t k2b_get_stat(k2b_handle* h, const char* name, double* value) {
  K2B_TRY(enter(h));
  if (name == nullptr || value == nullptr) return fail(h, K2B_ERR_INVALID, "k2b_get_stat: NULL argument");
  const std::string n(name);
  if (n == "decoder_table_bytes") *value = (double)h->dec_tab_bytes;
  else if (n == "decoder_table_build_ms") *value = (double)h->dec_tab_build_ms;
  else if (n == "decoder_table_state") *value = (double)h->dec_tab_state;
  else return fail(h, K2B_ERR_INVALID, "k2b_get_stat: unknown statistic " + n);
  return K2B_OK;
}

int32_t k2b_destroy(k2b_handle* h) {
  if (h == nullptr) return K2B_OK;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  float** ws[] = {&h->emb, &h->conv_w, &h->dec_w, &h->dec_b, &h->enc_w, &h->enc_b, &h->out_w, &h->out_b, &h->tab0, &h->tab1};
  for (float** p : ws) { if (*p) cudaFree(*p); *p = nullptr; }
  free_cluster_assets(h);
  state_pool_free(h);
  beam_pool_free(h);
  nccl_free(h);
  host_stage_free(h);
  if (h->lens_dev) cudaFree(h->lens_dev);
  if (h->cluster_timing) cudaFree(h->cluster_timing);
  if (h->timeline) cudaFree(h->timeline);
  if (h->dev_status) cudaFree(h->dev_status);
  for (int i = 0; i < 2; ++i) { if (h->ev_ready[i]) cudaEventDestroy(h->ev_ready[i]); if (h->ev_free[i]) cudaEventDestroy(h->ev_free[i]); }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->la_stream) {
    cudaStreamSynchronize(h->la_stream);
    cudaEventDestroy(h->la_ev_a); cudaEventDestroy(h->la_ev_b);
    for (int i = 0; i < 2; ++i) cudaEventDestroy(h->la_ev_done[i]);
    cudaEventDestroy(h->la_ev_mark);
    cudaStreamDestroy(h->la_stream);
  }
  free_buf(h->ws_encproj_la[0]); free_buf(h->ws_encproj_la[1]);
  if (h->la_flags) cudaFree(h->la_flags);
  if (h->cg_next) cudaFree(h->cg_next);
  if (h->cg_delta) cudaFree(h->cg_delta);
  if (h->cg_resid) cudaFree(h->cg_resid);
  free_buf(h->ws_cst);
  DevBuf* bufs[] = {&h->ws_in, &h->ws_encproj, &h->ws_x, &h->ws_ximg, &h->ws_dec, &h->ws_logits, &h->ws_part, &h->ws_state, &h->ws_bp,
                    &h->ws_out, &h->ws_misc, &h->ws_ctc, &h->ws_sync};
  for (DevBuf* b : bufs) free_buf(*b);
  for (cudaEvent_t ev : h->prof.start) cudaEventDestroy(ev);
  for (cudaEvent_t ev : h->prof.stop) cudaEventDestroy(ev);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return K2B_OK;
}

int32_t k2b_load_weights(k2b_handle* h, const float* emb, const float* conv_w, const float* dec_proj_w,
                         const float* dec_proj_b, const float* enc_proj_w, const float* enc_proj_b, const float* out_w,
                         const float* out_b) {
  K2B_TRY(enter(h));
  const k2b_config& c = h->cfg;
  if (!emb || !conv_w || !dec_proj_w || !dec_proj_b || !out_w || !out_b)
    return fail(h, K2B_ERR_INVALID, "k2b_load_weights: a required weight pointer is NULL");
  if (c.encoder_dim > 0 && (!enc_proj_w || !enc_proj_b))
    return fail(h, K2B_ERR_INVALID, "k2b_load_weights: encoder_dim > 0 needs enc_proj_w and enc_proj_b");
  const size_t V = c.vocab_size, J = c.joiner_dim, D = c.decoder_dim, E = c.encoder_dim, ctx = c.context_size;
  h->weights_loaded = false;
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  free_cluster_assets(h);
  K2B_TRY(upload(h, &h->emb, emb, V * D));
  K2B_TRY(upload(h, &h->conv_w, conv_w, D * 4 * ctx));
  K2B_TRY(upload(h, &h->dec_w, dec_proj_w, J * D));
  K2B_TRY(upload(h, &h->dec_b, dec_proj_b, J));
  K2B_TRY(upload(h, &h->out_w, out_w, V * J));
  K2B_TRY(upload(h, &h->out_b, out_b, V));
  if (E > 0) {
    K2B_TRY(upload(h, &h->enc_w, enc_proj_w, J * E));
    K2B_TRY(upload(h, &h->enc_b, enc_proj_b, J));
  }
  if (h->tab0) { cudaFree(h->tab0); h->tab0 = nullptr; }
  if (h->tab1) { cudaFree(h->tab1); h->tab1 = nullptr; }
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->tab0), (V + 1) * D * sizeof(float)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->tab1), (V + 1) * D * sizeof(float)));
  K2B_TRY(build_decoder_tables(h));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  h->weights_loaded = true;
  return K2B_OK;
}

int32_t k2b_set_precision(k2b_handle* h, int32_t precision) {
  K2B_TRY(enter(h));
  if (precision < K2B_PREC_FP32 || precision > K2B_PREC_BF16) return fail(h, K2B_ERR_INVALID, "unknown precision");
  if (precision != K2B_PREC_FP32 && h->cfg.joiner_dim % 64 != 0)
    return fail(h, K2B_ERR_UNSUPPORTED, "tcgen05 precisions need joiner_dim % 64 == 0");
  h->cfg.precision = precision;
  return K2B_OK;
}

int32_t k2b_set_stream(k2b_handle* h, void* cuda_stream) {
  K2B_TRY(enter(h));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return K2B_OK;
}

int32_t k2b_sync(k2b_handle* h) {
  K2B_TRY(enter(h));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  return cluster_status(h);          // device-pointer callers learn of kernel time-outs / rejected Hyp ids here
}

int64_t k2b_launch_count(const k2b_handle* h) { return h ? h->launches : 0; }

int32_t k2b_reset_launch_count(k2b_handle* h) {
  if (!h) return K2B_ERR_INVALID;
  h->launches = 0;
  return K2B_OK;
}

int32_t k2b_profile_enable(k2b_handle* h, int32_t on) {
  K2B_TRY(enter(h));
  h->profile_on = on != 0;
  return K2B_OK;
}

int32_t k2b_profile_read(k2b_handle* h, int64_t* n_launches, double* total_ms) {
  K2B_TRY(enter(h));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  ProfEvents& p = h->prof;
  for (size_t i = 0; i < p.used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.start[i], p.stop[i]) == cudaSuccess) { p.ms_total += ms; p.n_total++; }
  }
  p.used = 0;
  if (n_launches) *n_launches = p.n_total;
  if (total_ms) *total_ms = p.ms_total;
  p.n_total = 0;
  p.ms_total = 0.0;
  return K2B_OK;
}

// diagnostic: per-phase cycle totals of the last cluster-kernel launch (20 values); enables collection on first call,
// k2b_set_option("cluster_timing", 0) switches it off again (production launches then use the uninstrumented kernel)
K2B_API int32_t k2b_cluster_phase_cycles(k2b_handle* h, int64_t* out20) {
  K2B_TRY(enter(h));
  if (out20 == nullptr) return fail(h, K2B_ERR_INVALID, "k2b_cluster_phase_cycles: out20 is NULL");
  if (h->cluster_timing == nullptr) {
    K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->cluster_timing), 20 * sizeof(long long)));
    K2B_CUDA(h, cudaMemset(h->cluster_timing, 0, 20 * sizeof(long long)));
  }
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  long long v[20];
  K2B_CUDA(h, cudaMemcpy(v, h->cluster_timing, sizeof(v), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 20; ++i) out20[i] = v[i];
  K2B_CUDA(h, cudaMemset(h->cluster_timing, 0, sizeof(v)));
  return K2B_OK;
}

// diagnostic: the first call allocates the per-SM clock64 timeline of the per-frame beam path and switches collection on; later
// calls copy out [64][148][8] stamps (joiner: start, wait done, first accumulator, end; merge: first start, first wait done, last
// merge done, last end) and re-arm
K2B_API int32_t k2b_debug_timeline(k2b_handle* h, int64_t* out) {
  K2B_TRY(enter(h));
  const size_t n = (size_t)64 * 148 * 8;
  std::vector<long long> init(n);
  for (size_t i = 0; i < n; ++i) init[i] = ((i & 7) == 4 || (i & 7) == 5) ? 0x7fffffffffffffffll : 0ll;
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->timeline == nullptr) {
    K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->timeline), n * sizeof(long long)));
  } else if (out != nullptr) {
    K2B_CUDA(h, cudaMemcpy(out, h->timeline, n * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  K2B_CUDA(h, cudaMemcpy(h->timeline, init.data(), n * sizeof(long long), cudaMemcpyHostToDevice));
  h->timeline_frame = 0;
  return K2B_OK;
}

// ---- fine-grained -------------------------------------------------------------------------------
int32_t k2b_decoder_proj_dev(k2b_handle* h, const int64_t* y, int32_t n, float* out) {
  K2B_TRY(enter(h));
  K2B_TRY(need_weights(h));
  if (n < 0 || (n > 0 && out == nullptr)) return fail(h, K2B_ERR_INVALID, "k2b_decoder_proj: bad n or out");
  if (n == 0) return K2B_OK;
  K2B_TRY(ensure(h, h->ws_misc, sizeof(int32_t) * 2 * (size_t)n));
  int32_t* ctx = static_cast<int32_t*>(h->ws_misc.p);
  ctx_from_i64_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(y, n, h->cfg.blank_id, ctx);
  K2B_LAUNCH_CHECK(h);
  GemmArgs a;
  a.M = n; a.N = h->cfg.joiner_dim; a.K = h->cfg.decoder_dim;
  a.W = h->dec_w; a.bias = h->dec_b;
  a.ctx = ctx; a.tab0 = h->tab0; a.tab1 = h->tab1; a.V = h->cfg.vocab_size;
  a.neg_wrap = h->cfg.neg_id_mode == K2B_NEGID_WRAP; a.blank = h->cfg.blank_id;
  a.C = out;
  return launch_gemm_simt(h, PRO_DEC, EPI_STORE, a);
}

int32_t k2b_decoder_proj(k2b_handle* h, const int64_t* y, int32_t n, float* out) {
  K2B_TRY(enter(h));
  K2B_TRY(need_weights(h));
  if (n < 0 || (n > 0 && out == nullptr)) return fail(h, K2B_ERR_INVALID, "k2b_decoder_proj: bad n or out");
  if (n == 0) return K2B_OK;
  const size_t J = h->cfg.joiner_dim;
  K2B_TRY(ensure(h, h->ws_dec, sizeof(float) * (size_t)n * J));
  const int64_t* yd = nullptr;
  if (y != nullptr) {
    K2B_TRY(ensure(h, h->ws_in, sizeof(int64_t) * 2 * (size_t)n));
    K2B_CUDA(h, cudaMemcpyAsync(h->ws_in.p, y, sizeof(int64_t) * 2 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    yd = static_cast<const int64_t*>(h->ws_in.p);
  }
  K2B_TRY(k2b_decoder_proj_dev(h, yd, n, static_cast<float*>(h->ws_dec.p)));
  K2B_CUDA(h, cudaMemcpyAsync(out, h->ws_dec.p, sizeof(float) * (size_t)n * J, cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  return K2B_OK;
}

int32_t k2b_joiner_proj_dev(k2b_handle* h, const float* enc, const float* dec, int32_t n, float* logits) {
  K2B_TRY(enter(h));
  K2B_TRY(need_weights(h));
  if (n < 0 || (n > 0 && (!enc || !dec || !logits))) return fail(h, K2B_ERR_INVALID, "k2b_joiner_proj: bad argument");
  if (n == 0) return K2B_OK;
  GemmArgs a;
  a.M = n; a.N = h->cfg.vocab_size; a.K = h->cfg.joiner_dim;
  a.W = h->out_w; a.bias = h->out_b;
  a.enc = enc; a.enc_stride = h->cfg.joiner_dim; a.rows_per_stream = 1; a.dec = dec;
  a.C = logits;
  prof_begin(h);
  int32_t s = launch_gemm_simt(h, PRO_JOIN, EPI_STORE, a);
  prof_end(h);
  return s;
}

int32_t k2b_joiner_proj(k2b_handle* h, const float* enc, const float* dec, int32_t n, float* logits) {
  K2B_TRY(enter(h));
  K2B_TRY(need_weights(h));
  if (n < 0 || (n > 0 && (!enc || !dec || !logits))) return fail(h, K2B_ERR_INVALID, "k2b_joiner_proj: bad argument");
  if (n == 0) return K2B_OK;
  const size_t J = h->cfg.joiner_dim, V = h->cfg.vocab_size;
  K2B_TRY(ensure(h, h->ws_in, sizeof(float) * 2 * (size_t)n * J));
  K2B_TRY(ensure(h, h->ws_logits, sizeof(float) * (size_t)n * V));
  float* de = static_cast<float*>(h->ws_in.p);
  float* dd = de + (size_t)n * J;
  K2B_CUDA(h, cudaMemcpyAsync(de, enc, sizeof(float) * (size_t)n * J, cudaMemcpyHostToDevice, h->stream));
  K2B_CUDA(h, cudaMemcpyAsync(dd, dec, sizeof(float) * (size_t)n * J, cudaMemcpyHostToDevice, h->stream));
  K2B_TRY(k2b_joiner_proj_dev(h, de, dd, n, static_cast<float*>(h->ws_logits.p)));
  K2B_CUDA(h, cudaMemcpyAsync(logits, h->ws_logits.p, sizeof(float) * (size_t)n * V, cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  return K2B_OK;
}

int32_t k2b_encoder_proj_dev(k2b_handle* h, const float* raw, int32_t n, float* out) {
  K2B_TRY(enter(h));
  K2B_TRY(need_weights(h));
  if (n < 0 || (n > 0 && (!raw || !out))) return fail(h, K2B_ERR_INVALID, "k2b_encoder_proj: bad argument");
  if (n == 0) return K2B_OK;
  return encoder_proj_launch(h, raw, n, out);
}

int32_t k2b_encoder_proj(k2b_handle* h, const float* raw, int32_t n, float* out) {
  K2B_TRY(enter(h));
  K2B_TRY(need_weights(h));
  if (n < 0 || (n > 0 && (!raw || !out))) return fail(h, K2B_ERR_INVALID, "k2b_encoder_proj: bad argument");
  if (n == 0) return K2B_OK;
  const size_t J = h->cfg.joiner_dim, E = h->cfg.encoder_dim;
  K2B_TRY(ensure(h, h->ws_in, sizeof(float) * (size_t)n * E));
  K2B_TRY(ensure(h, h->ws_encproj, sizeof(float) * (size_t)n * J));
  K2B_CUDA(h, cudaMemcpyAsync(h->ws_in.p, raw, sizeof(float) * (size_t)n * E, cudaMemcpyHostToDevice, h->stream));
  K2B_TRY(encoder_proj_launch(h, static_cast<const float*>(h->ws_in.p), n, static_cast<float*>(h->ws_encproj.p)));
  K2B_CUDA(h, cudaMemcpyAsync(out, h->ws_encproj.p, sizeof(float) * (size_t)n * J, cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->cfg.precision != K2B_PREC_FP32) K2B_TRY(cluster_status(h));
  return K2B_OK;
}

// ---- on-device streaming state ---------------------------------------------------------------------------
int32_t k2b_state_pool_create(k2b_handle* h, const int32_t* item_len, int32_t n_tensors, int32_t max_streams) {
  K2B_TRY(enter(h));
  if (item_len == nullptr || n_tensors <= 0 || max_streams <= 0) return fail(h, K2B_ERR_INVALID, "k2b_state_pool_create: bad arguments");
  return state_pool_create(h, item_len, n_tensors, max_streams);
}
int64_t k2b_state_pool_stacked_floats(k2b_handle* h, int32_t B) { return h == nullptr ? 0 : (int64_t)state_pool_stacked_floats(h, B); }
int32_t k2b_state_pool_put(k2b_handle* h, int32_t slot, const float* state) {
  K2B_TRY(enter(h));
  if (state == nullptr) return fail(h, K2B_ERR_INVALID, "k2b_state_pool_put: state is NULL");
  return state_pool_io(h, slot, const_cast<float*>(state), true);
}
int32_t k2b_state_pool_get(k2b_handle* h, int32_t slot, float* state) {
  K2B_TRY(enter(h));
  if (state == nullptr) return fail(h, K2B_ERR_INVALID, "k2b_state_pool_get: state is NULL");
  return state_pool_io(h, slot, state, false);
}
int32_t k2b_stack_states(k2b_handle* h, const int32_t* slots, int32_t B, const int32_t* axis_len, float* stacked) {
  K2B_TRY(enter(h));
  if (B < 0 || (B > 0 && (slots == nullptr || axis_len == nullptr || stacked == nullptr))) return fail(h, K2B_ERR_INVALID, "k2b_stack_states: bad arguments");
  return state_pool_restack(h, slots, B, axis_len, stacked, false);
}
int32_t k2b_unstack_states(k2b_handle* h, const int32_t* slots, int32_t B, const int32_t* axis_len, const float* stacked) {
  K2B_TRY(enter(h));
  if (B < 0 || (B > 0 && (slots == nullptr || axis_len == nullptr || stacked == nullptr))) return fail(h, K2B_ERR_INVALID, "k2b_unstack_states: bad arguments");
  return state_pool_restack(h, slots, B, axis_len, const_cast<float*>(stacked), true);
}

// ---- ragged batches -----------------------------------------------------------------------------------
// The per-stream frame counts set by k2b_set_encoder_out_lens are consumed by the next fused offline search call; an online chunk
// call in between discards them (the online loops have no lengths: ref OnlineRecognizer.cs:85-219).
namespace {
struct LensGuard {
  k2b_handle* h;
  explicit LensGuard(k2b_handle* hh) : h(hh) {}
  ~LensGuard() { if (h != nullptr) h->lens_active = false; }
};
// Host-pointer wrappers stage their input with a copy on the handle's stream and then run the device-pointer code: whatever the
// caller promised about ITS device buffers ("inputs_complete"), the staged frames are only ordered on the handle's stream.
struct StagedInputGuard {
  k2b_handle* h;
  int keep;
  explicit StagedInputGuard(k2b_handle* hh) : h(hh), keep(hh->opt_inputs_complete) { h->opt_inputs_complete = 0; }
  ~StagedInputGuard() { h->opt_inputs_complete = keep; }
};
int32_t check_lens(k2b_handle* h, int B, const char* who) {
  if (h->lens_active && h->lens_n != B) {
    return fail(h, K2B_ERR_INVALID, std::string(who) + ": k2b_set_encoder_out_lens was given " + std::to_string(h->lens_n) +
                                        " streams, this call has " + std::to_string(B));
  }
  return K2B_OK;
}

int32_t check_hyp_host(k2b_handle* h, const int64_t* hyp, int B, bool allow_neg_tail, const char* who) {
  const int64_t V = h->cfg.vocab_size;
  for (int b = 0; b < B; ++b) {
    const int64_t c0 = hyp[2 * b], c1 = hyp[2 * b + 1];
    if (c0 < -1 || c0 >= V || c1 >= V || c1 < (allow_neg_tail ? -1 : 0))
      return fail(h, K2B_ERR_INVALID, std::string(who) + ": Hyp of stream " + std::to_string(b) + " holds a token id outside the vocabulary");
  }
  return K2B_OK;
}

// host-pointer fused calls end here: results are on their way to the caller's buffers; wait for them (and for the kernels' status
// word) unless the caller asked for overlap (k2b_set_option("async_d2h", 1) with pinned buffers: k2b_sync completes the call)
int32_t finish_host_call(k2b_handle* h) {
  if (h->opt_async_d2h) return K2B_OK;
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  return cluster_status(h);
}

}  // namespace

int32_t k2b_set_encoder_out_lens(k2b_handle* h, const int64_t* lens, int32_t B) {
  K2B_TRY(enter(h));
  h->lens_active = false;
  if (lens == nullptr || B <= 0) return K2B_OK;
  std::vector<int32_t> v((size_t)B);
  for (int32_t b = 0; b < B; ++b) {
    if (lens[b] < 0 || lens[b] > 0x7fffffff) return fail(h, K2B_ERR_INVALID, "k2b_set_encoder_out_lens: negative or oversized length");
    v[(size_t)b] = (int32_t)lens[b];
  }
  if (h->lens_n < B || h->lens_dev == nullptr) {
    if (h->lens_dev) cudaFree(h->lens_dev);
    h->lens_dev = nullptr;
    K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->lens_dev), sizeof(int32_t) * (size_t)B));
  }
  // synchronous on purpose: `v` is pageable and goes out of scope; earlier searches that still read the buffer finish first
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  K2B_CUDA(h, cudaMemcpy(h->lens_dev, v.data(), sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice));
  h->lens_n = B;
  h->lens_active = true;
  return K2B_OK;
}

// Greedy search runs as beam 1: on the cluster kernel wherever a cluster holds the vocabulary (V <= 2048 with sixteen-CTA clusters),
// on the persistent beam kernel beyond that - and also for 1024 < V <= 2048 when the batch needs three or more waves of 16-CTA
// clusters (about six of them, 32 streams each, are co-resident). Measured on cfg3 (V = 2000, 512 streams, chunks of 8 frames):
// cluster kernel 185 us per chunk, persistent kernel 147 us (its merge warps step four streams at once, greedy_merge_warp).
// k2b_set_option("greedy_persistent", 0/1) overrides the choice (comparison runs, tests).
static bool prefer_persistent_greedy(k2b_handle* h, int B) {
  if (h->opt_greedy_persistent >= 0) return h->opt_greedy_persistent != 0 && beam_greedy_usable(h);
  return h->cfg.vocab_size > 1024 && B > 384 && beam_greedy_usable(h);
}

// One pass of per-stream greedy search on the fast engines (beam 1 on the cluster kernel or on the persistent kernel) over frames
// [t_begin, T) of a [B,T,.] batch. hyp_in (device, [B,2]) seeds the contexts (nullptr = {-1, blank}); extra_mask is the online
// loop's literal 1 or -1. `frames_ready`: ws_encproj already holds this batch's frames in the engine's form (second pass of
// BATCH_COMPAT). Timestamps are frame indices of the whole utterance on the cluster kernel and relative to t_begin on the
// persistent kernel (*ts_rel tells).
static int32_t greedy_fast_pass(k2b_handle* h, const float* enc, int enc_is_raw, int B, int T, int t_begin, int extra_mask,
                                int64_t* hyp_inout, int64_t* tokens, int32_t* ts, int32_t* n_out, int cap, bool frames_ready,
                                bool* ts_rel) {
  const size_t J = h->cfg.joiner_dim;
  *ts_rel = false;
  if (cluster_path_supported(h, 1) && !prefer_persistent_greedy(h, B)) {
    K2B_TRY(ensure_cluster_assets(h));
    const size_t n = (size_t)B * T;
    K2B_TRY(ensure(h, h->ws_encproj, sizeof(float) * n * J));
    float* encE = static_cast<float*>(h->ws_encproj.p);
    if (!frames_ready) {
      if (enc_is_raw) {
        if (h->cfg.encoder_dim <= 0 || h->enc_w == nullptr) return fail(h, K2B_ERR_STATE, "encoder_proj weights not loaded (E == 0)");
        if (encproj_tc_supported(h)) {
          K2B_TRY(encoder_proj_tc(h, enc, (int)n, encE, true));
        } else {
          GemmArgs g;
          g.M = (int)n; g.N = (int)J; g.K = h->cfg.encoder_dim;
          g.A = enc; g.W = h->enc_w; g.bias = h->enc_b; g.C = encE;
          K2B_TRY(launch_gemm_simt(h, PRO_PLAIN, EPI_EXP2X, g));
        }
      } else {
        K2B_TRY(exp2x_frames(h, enc, encE, n * J));
      }
    }
    // a few streams: one 8-CTA cluster per stream, weights in registers, fp32 FMA - no back-pointers, no back-trace
    if (single_greedy_usable(h, B, T - t_begin))
      return single_greedy_dev(h, encE, B, T - t_begin, t_begin, T, extra_mask, hyp_inout, hyp_inout, tokens, ts, n_out, cap);
    const size_t a4 = ((size_t)B * 4 + 255) & ~size_t(255);
    K2B_TRY(ensure(h, h->ws_state, 4 * a4));
    K2B_TRY(ensure(h, h->ws_bp, sizeof(int32_t) * (size_t)B * T));
    char* p = static_cast<char*>(h->ws_state.p);
    float* fin_lp = reinterpret_cast<float*>(p); p += a4;
    int32_t* fin_len = reinterpret_cast<int32_t*>(p); p += a4;
    int32_t* fin_nlive = reinterpret_cast<int32_t*>(p); p += a4;
    float* score = reinterpret_cast<float*>(p);
    int32_t* bp = static_cast<int32_t*>(h->ws_bp.p);
    if (t_begin > 0) K2B_CUDA(h, cudaMemsetAsync(bp, 0, sizeof(int32_t) * (size_t)B * T, h->stream));   // rows before t_begin: "nothing emitted"
    K2B_TRY(beam_cluster_dev(h, encE, B, T - t_begin, 1, bp, fin_lp, fin_len, fin_nlive, extra_mask, hyp_inout, hyp_inout, t_begin, T, 0,
                             nullptr, nullptr, false));
    return beam_backtrace_dev(h, B, 1, T, fin_lp, fin_len, fin_nlive, bp, tokens, ts, n_out, score, cap);
  }
  // large vocabulary: beam 1 on the persistent beam kernel
  const float* frames = static_cast<const float*>(h->ws_encproj.p);
  // "inputs_complete" (online chunks that follow each other, 128 < B <= 4 * (SMs - 20) so that the merges still are one wave): the
  // raw frames are projected on the low-priority side stream, which is not ordered behind the handle's stream, into one of two
  // buffers taken in turns, and the persistent kernel leaves 20 SMs free - so the projection of chunk c+1 runs under the search
  // of chunk c (cfg3: 19 of the 114 us per chunk)
  const int cap_sms = h->sm_count - 20;
  const bool ahead = !frames_ready && enc_is_raw && h->opt_inputs_complete != 0 && t_begin == 0 && encproj_tc_supported(h) &&
                     h->cfg.encoder_dim > 0 && h->enc_w != nullptr && B > 128 && B <= 4 * cap_sms && !h->profile_on;
  if (ahead) {
    if (h->la_stream == nullptr) K2B_TRY(lookahead_init(h));
    const int X = (int)(h->la_turn++ & 1u);
    DevBuf& buf = h->ws_encproj_la[X];
    const size_t need = sizeof(float) * (size_t)B * T * J;
    if (need > buf.bytes) {
      K2B_CUDA(h, cudaStreamSynchronize(h->la_stream));
      K2B_TRY(ensure(h, buf, need));
      h->la_done_valid[X] = false;
    }
    const bool cold = !h->enc_ready;
    K2B_TRY(ensure_encproj_assets(h));
    if (cold) K2B_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->la_done_valid[X]) K2B_CUDA(h, cudaStreamWaitEvent(h->la_stream, h->la_ev_done[X], 0));
    if (h->la_mark_valid) K2B_CUDA(h, cudaStreamWaitEvent(h->la_stream, h->la_ev_mark, 0));
    cudaStream_t keep = h->stream;
    h->stream = h->la_stream;
    const int32_t st = encoder_proj_tc(h, enc, B * T, static_cast<float*>(buf.p), false);
    h->stream = keep;
    K2B_TRY(st);
    K2B_CUDA(h, cudaEventRecord(h->la_ev_a, h->la_stream));
    K2B_CUDA(h, cudaStreamWaitEvent(h->stream, h->la_ev_a, 0));
    K2B_TRY(ensure(h, h->ws_misc, sizeof(float) * (size_t)B));
    h->mega_grid_cap = cap_sms;
    h->la_mark_arm = true;
    const int32_t sb = beam_dev(h, static_cast<const float*>(buf.p), B, T, 1, tokens, ts, n_out, static_cast<float*>(h->ws_misc.p), cap,
                                extra_mask, hyp_inout, true, 0, 0, false, 0);
    h->mega_grid_cap = 0;
    h->la_mark_arm = false;
    K2B_TRY(sb);
    K2B_CUDA(h, cudaEventRecord(h->la_ev_done[X], h->stream));
    h->la_done_valid[X] = true;
    return K2B_OK;
  }
  if (!frames_ready) K2B_TRY(frames_for_search(h, enc, enc_is_raw, B, T, &frames));
  else if (!enc_is_raw) frames = enc;
  K2B_TRY(ensure(h, h->ws_misc, sizeof(float) * (size_t)B));
  *ts_rel = t_begin > 0;
  return beam_dev(h, frames + (size_t)t_begin * J, B, T - t_begin, 1, tokens, ts, n_out, static_cast<float*>(h->ws_misc.p), cap, extra_mask,
                  hyp_inout, true, 0, 0, false, t_begin > 0 ? (long long)T * (long long)J : 0);
}

// BATCH_COMPAT on the fast engines. The reference's batch loop (ref OfflineRecognizer.cs:189-303) re-runs the decoder for ALL
// streams from their token-list tails whenever ANY stream emits (ref :246, :278-286); the lists are seeded with blanks (Q5), so
// the only effect on a stream is that its context is {blank, blank} instead of {-1, blank} from the frame after the batch's
// first emission (t*) until its own first emission (Q6) - from then on the tail of its list is what the single-stream loop
// would use. Hence: pass 1 decodes every stream alone (PER_STREAM); t* = min of the first timestamps; streams that emit at t*
// are done; the others emitted nothing up to t* and are decoded again over (t*, T) from {blank, blank} (pass 2).
static int32_t greedy_compat_fast(k2b_handle* h, const float* enc, int enc_is_raw, int B, int T, int64_t* tokens, int32_t* ts,
                                  int32_t* n_out, int cap) {
  bool rel = false;
  K2B_TRY(greedy_fast_pass(h, enc, enc_is_raw, B, T, 0, -1, nullptr, tokens, ts, n_out, cap, false, &rel));
  int* tstar_dev = h->dev_status + 3;
  int tstar = T;
  K2B_CUDA(h, cudaMemcpyAsync(tstar_dev, &tstar, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  first_emission_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(ts, n_out, B, cap, T, tstar_dev);
  K2B_LAUNCH_CHECK(h);
  K2B_CUDA(h, cudaMemcpyAsync(&tstar, tstar_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));      // the second pass starts at a data-dependent frame
  if (tstar + 1 >= T) return K2B_OK;
  const size_t tok_b = (sizeof(int64_t) * (size_t)B * cap + 255) & ~size_t(255), ts_b = (sizeof(int32_t) * (size_t)B * cap + 255) & ~size_t(255);
  const size_t n_b = (sizeof(int32_t) * (size_t)B + 255) & ~size_t(255), hyp_b = (sizeof(int64_t) * 2 * (size_t)B + 255) & ~size_t(255);
  K2B_TRY(ensure(h, h->ws_dec, tok_b + ts_b + n_b + hyp_b));
  char* p = static_cast<char*>(h->ws_dec.p);
  int64_t* tok2 = reinterpret_cast<int64_t*>(p); p += tok_b;
  int32_t* ts2 = reinterpret_cast<int32_t*>(p); p += ts_b;
  int32_t* n2 = reinterpret_cast<int32_t*>(p); p += n_b;
  int64_t* hyp2 = reinterpret_cast<int64_t*>(p);
  fill_hyp_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(hyp2, B, h->cfg.blank_id, h->cfg.blank_id);
  K2B_LAUNCH_CHECK(h);
  K2B_TRY(greedy_fast_pass(h, enc, enc_is_raw, B, T, tstar + 1, -1, hyp2, tok2, ts2, n2, cap, true, &rel));
  compat_select_kernel<<<B, 64, 0, h->stream>>>(B, cap, tstar, rel ? tstar + 1 : 0, tok2, ts2, n2, tokens, ts, n_out);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

// ---- fused search ---------------------------------------------------------------------------------
int32_t k2b_greedy_offline_dev(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t T, int32_t mode,
                               int64_t* tokens, int32_t* ts, int32_t* n_out, int32_t cap) {
  K2B_TRY(enter(h));
  K2B_TRY(need_weights(h));
  LensGuard lens_guard(h);
  K2B_TRY(check_search_args(h, enc, B, T, cap, tokens, ts, n_out, "k2b_greedy_offline"));
  K2B_TRY(check_lens(h, B, "k2b_greedy_offline"));
  if (mode < K2B_GREEDY_SINGLE || mode > K2B_GREEDY_PER_STREAM) return fail(h, K2B_ERR_INVALID, "k2b_greedy_offline: unknown mode");
  if (h->lens_active && mode == K2B_GREEDY_BATCH_COMPAT)
    return fail(h, K2B_ERR_UNSUPPORTED, "k2b_greedy_offline: BATCH_COMPAT decodes the padding like the reference (Q7); use PER_STREAM with lengths");
  if (mode == K2B_GREEDY_SINGLE && B != 1) return fail(h, K2B_ERR_INVALID, "k2b_greedy_offline: SINGLE mode needs B == 1");
  if (B == 0) return K2B_OK;
  // greedy search == beam 1 with the same tie rule: SINGLE / PER_STREAM run on the persistent kernels in the tcgen05 precisions,
  // BATCH_COMPAT as two such passes (greedy_compat_fast). The per-frame path keeps: fp32 precision, an utterance long enough to
  // meet the 1000-symbol cap of the single-stream loop (ref OfflineRecognizer.cs:122), more than one symbol per frame.
  const bool fast = h->cfg.precision != K2B_PREC_FP32 && T > 0 && T <= 1000 && h->max_sym_per_frame == 1 &&
                    (cluster_path_supported(h, 1) || beam_greedy_usable(h));
  if (fast && mode == K2B_GREEDY_BATCH_COMPAT && B > 1) return greedy_compat_fast(h, enc, enc_is_raw, B, T, tokens, ts, n_out, cap);
  if (fast) {
    bool rel = false;
    return greedy_fast_pass(h, enc, enc_is_raw, B, T, 0, -1, nullptr, tokens, ts, n_out, cap, false, &rel);
  }
  const float* frames = nullptr;
  K2B_TRY(frames_for_search(h, enc, enc_is_raw, B, T, &frames));
  return greedy_dev(h, frames, B, T, mode, false, nullptr, tokens, ts, n_out, cap);
}

int32_t k2b_greedy_offline(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t T, int32_t mode,
                           int64_t* tokens, int32_t* ts, int32_t* n_out, int32_t cap) {
  K2B_TRY(enter(h));
  StagedInputGuard staged_guard(h);
  LensGuard lens_guard(h);
  K2B_TRY(need_weights(h));
  K2B_TRY(check_search_args(h, enc, B, T, cap, tokens, ts, n_out, "k2b_greedy_offline"));
  if (B == 0) return K2B_OK;
  const size_t width = enc_is_raw ? h->cfg.encoder_dim : h->cfg.joiner_dim;
  const size_t in_bytes = sizeof(float) * (size_t)B * T * width;
  K2B_TRY(ensure(h, h->ws_in, in_bytes > 0 ? in_bytes : 4));
  OutStage o;
  K2B_TRY(stage_out(h, B, cap > 0 ? cap : 1, &o));
  if (in_bytes) K2B_TRY(h2d_rows(h, h->ws_in.p, in_bytes, enc, in_bytes, in_bytes, 1, h->stream));
  K2B_TRY(k2b_greedy_offline_dev(h, static_cast<const float*>(h->ws_in.p), enc_is_raw, B, T, mode, o.tokens, o.ts, o.n, cap));
  if (cap > 0) {
    K2B_CUDA(h, cudaMemcpyAsync(tokens, o.tokens, sizeof(int64_t) * (size_t)B * cap, cudaMemcpyDeviceToHost, h->stream));
    K2B_CUDA(h, cudaMemcpyAsync(ts, o.ts, sizeof(int32_t) * (size_t)B * cap, cudaMemcpyDeviceToHost, h->stream));
  }
  K2B_CUDA(h, cudaMemcpyAsync(n_out, o.n, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  return finish_host_call(h);
}

int32_t k2b_greedy_online_chunk_dev(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t Tc,
                                    int64_t* hyp_inout, int64_t* tokens, int32_t* ts, int32_t* n_out, int32_t cap) {
  K2B_TRY(enter(h));
  K2B_TRY(need_weights(h));
  h->lens_active = false;              // lengths belong to offline searches; a pending setting must not freeze online streams
  K2B_TRY(check_search_args(h, enc, B, Tc, cap, tokens, ts, n_out, "k2b_greedy_online_chunk"));
  if (B > 0 && hyp_inout == nullptr) return fail(h, K2B_ERR_INVALID, "k2b_greedy_online_chunk: hyp_inout is NULL");
  if (B == 0) return K2B_OK;
  const bool as_beam1 = h->cfg.precision != K2B_PREC_FP32 && Tc > 0 &&                   // online: Q6 is a no-op (ctx == list tail)
                        (cluster_path_supported(h, 1) || beam_greedy_usable(h));
  check_hyp_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(hyp_inout, B, h->cfg.vocab_size, h->cfg.blank_id, as_beam1 ? 0 : 1, h->dev_status);
  K2B_LAUNCH_CHECK(h);
  if (as_beam1) {
    bool rel = false;
    return greedy_fast_pass(h, enc, enc_is_raw, B, Tc, 0, 1, hyp_inout, tokens, ts, n_out, cap, false, &rel);
  }
  const float* frames = nullptr;
  K2B_TRY(frames_for_search(h, enc, enc_is_raw, B, Tc, &frames));
  return greedy_dev(h, frames, B, Tc, K2B_GREEDY_BATCH_COMPAT, true, hyp_inout, tokens, ts, n_out, cap);
}

int32_t k2b_greedy_online_chunk(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t Tc,
                                int64_t* hyp_inout, int64_t* tokens, int32_t* ts, int32_t* n_out, int32_t cap) {
  K2B_TRY(enter(h));
  StagedInputGuard staged_guard(h);
  K2B_TRY(need_weights(h));
  K2B_TRY(check_search_args(h, enc, B, Tc, cap, tokens, ts, n_out, "k2b_greedy_online_chunk"));
  if (B > 0 && hyp_inout == nullptr) return fail(h, K2B_ERR_INVALID, "k2b_greedy_online_chunk: hyp_inout is NULL");
  if (B == 0) return K2B_OK;
  K2B_TRY(check_hyp_host(h, hyp_inout, B, h->cfg.precision == K2B_PREC_FP32, "k2b_greedy_online_chunk"));
  const size_t width = enc_is_raw ? h->cfg.encoder_dim : h->cfg.joiner_dim;
  const size_t in_bytes = sizeof(float) * (size_t)B * Tc * width;
  K2B_TRY(ensure(h, h->ws_in, in_bytes > 0 ? in_bytes : 4));
  OutStage o;
  K2B_TRY(stage_out(h, B, cap > 0 ? cap : 1, &o));
  if (in_bytes) K2B_TRY(h2d_rows(h, h->ws_in.p, in_bytes, enc, in_bytes, in_bytes, 1, h->stream));
  K2B_CUDA(h, cudaMemcpyAsync(o.hyp, hyp_inout, sizeof(int64_t) * 2 * (size_t)B, cudaMemcpyHostToDevice, h->stream));
  K2B_TRY(k2b_greedy_online_chunk_dev(h, static_cast<const float*>(h->ws_in.p), enc_is_raw, B, Tc, o.hyp, o.tokens, o.ts, o.n, cap));
  if (cap > 0) {
    K2B_CUDA(h, cudaMemcpyAsync(tokens, o.tokens, sizeof(int64_t) * (size_t)B * cap, cudaMemcpyDeviceToHost, h->stream));
    K2B_CUDA(h, cudaMemcpyAsync(ts, o.ts, sizeof(int32_t) * (size_t)B * cap, cudaMemcpyDeviceToHost, h->stream));
  }
  K2B_CUDA(h, cudaMemcpyAsync(n_out, o.n, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaMemcpyAsync(hyp_inout, o.hyp, sizeof(int64_t) * 2 * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  return finish_host_call(h);
}

int32_t k2b_modified_beam_search_dev(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t T, int32_t K,
                                     int64_t* tokens, int32_t* ts, int32_t* n_out, float* score, int32_t cap) {
  K2B_TRY(enter(h, false));
  K2B_TRY(need_weights(h));
  LensGuard lens_guard(h);
  K2B_TRY(check_search_args(h, enc, B, T, cap, tokens, ts, n_out, "k2b_modified_beam_search"));
  K2B_TRY(check_lens(h, B, "k2b_modified_beam_search"));
  if (K < 1 || K > kMaxBeam) return fail(h, K2B_ERR_INVALID, "k2b_modified_beam_search: beam must be in 1..8");
  if (K > h->cfg.vocab_size) return fail(h, K2B_ERR_INVALID, "k2b_modified_beam_search: beam exceeds vocab_size");
  if (B > 0 && score == nullptr) return fail(h, K2B_ERR_INVALID, "k2b_modified_beam_search: score is NULL");
  if (B == 0) return K2B_OK;
  if (h->cfg.precision != K2B_PREC_FP32 && cluster_path_supported(h, K) && T > 0)
    return beam_cluster_path(h, enc, enc_is_raw, B, T, K, tokens, ts, n_out, score, cap);   // joins in front of its back-trace
  K2B_TRY(gather_join(h));
  const float* frames = nullptr;
  K2B_TRY(frames_for_search(h, enc, enc_is_raw, B, T, &frames));
  return beam_dev(h, frames, B, T, K, tokens, ts, n_out, score, cap);
}

int32_t k2b_modified_beam_search(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t T, int32_t K,
                                 int64_t* tokens, int32_t* ts, int32_t* n_out, float* score, int32_t cap) {
  K2B_TRY(enter(h));
  StagedInputGuard staged_guard(h);
  LensGuard lens_guard(h);
  K2B_TRY(need_weights(h));
  K2B_TRY(check_search_args(h, enc, B, T, cap, tokens, ts, n_out, "k2b_modified_beam_search"));
  K2B_TRY(check_lens(h, B, "k2b_modified_beam_search"));
  if (B > 0 && score == nullptr) return fail(h, K2B_ERR_INVALID, "k2b_modified_beam_search: score is NULL");
  if (B == 0) return K2B_OK;
  const size_t width = enc_is_raw ? h->cfg.encoder_dim : h->cfg.joiner_dim;
  const size_t in_bytes = sizeof(float) * (size_t)B * T * width;
  OutStage o;
  K2B_TRY(stage_out(h, B, cap > 0 ? cap : 1, &o));
  if (K < 1 || K > kMaxBeam) return fail(h, K2B_ERR_INVALID, "k2b_modified_beam_search: beam must be in 1..8");
  const bool pipelined = h->cfg.precision != K2B_PREC_FP32 && cluster_path_supported(h, K) && (!enc_is_raw || encproj_tc_supported(h)) &&
                         T >= 32 && !h->profile_on && K <= h->cfg.vocab_size && h->cfg.joiner_dim % 4 == 0;
  const bool chunked = !pipelined && h->cfg.precision != K2B_PREC_FP32 && !cluster_path_supported(h, K) &&
                       T >= 64 && !h->profile_on && K <= h->cfg.vocab_size && (!enc_is_raw || h->cfg.encoder_dim > 0) && beam_chunkable(h, K);
  if (pipelined) {
    K2B_TRY(beam_cluster_pipelined(h, enc, enc_is_raw, B, T, K, o.tokens, o.ts, o.n, o.score, cap));
  } else if (chunked) {
    K2B_TRY(beam_chunked_host(h, enc, enc_is_raw, B, T, K, o.tokens, o.ts, o.n, o.score, cap));
  } else {
    K2B_TRY(ensure(h, h->ws_in, in_bytes > 0 ? in_bytes : 4));
    if (in_bytes) K2B_TRY(h2d_rows(h, h->ws_in.p, in_bytes, enc, in_bytes, in_bytes, 1, h->stream));
    K2B_TRY(k2b_modified_beam_search_dev(h, static_cast<const float*>(h->ws_in.p), enc_is_raw, B, T, K, o.tokens, o.ts, o.n,
                                         o.score, cap));
  }
  if (cap > 0) {
    K2B_CUDA(h, cudaMemcpyAsync(tokens, o.tokens, sizeof(int64_t) * (size_t)B * cap, cudaMemcpyDeviceToHost, h->stream));
    K2B_CUDA(h, cudaMemcpyAsync(ts, o.ts, sizeof(int32_t) * (size_t)B * cap, cudaMemcpyDeviceToHost, h->stream));
  }
  K2B_CUDA(h, cudaMemcpyAsync(n_out, o.n, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaMemcpyAsync(score, o.score, sizeof(float) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  return finish_host_call(h);
}

// ---- contextual biasing (hot words) -----------------------------------------------------------------------------------
int32_t k2b_set_context_graph(k2b_handle* h, const int32_t* next, const float* delta, const float* residual, int32_t n_states) {
  K2B_TRY(enter(h));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->cg_next) cudaFree(h->cg_next);
  if (h->cg_delta) cudaFree(h->cg_delta);
  if (h->cg_resid) cudaFree(h->cg_resid);
  h->cg_next = nullptr; h->cg_delta = nullptr; h->cg_resid = nullptr; h->cg_states = 0;
  if (n_states <= 0 || next == nullptr) return K2B_OK;                       // cleared
  if (delta == nullptr || residual == nullptr) return fail(h, K2B_ERR_INVALID, "k2b_set_context_graph: delta / residual is NULL");
  const size_t V = h->cfg.vocab_size, n = (size_t)n_states * V;
  for (size_t i = 0; i < n; ++i)
    if (next[i] < 0 || next[i] >= n_states) return fail(h, K2B_ERR_INVALID, "k2b_set_context_graph: a transition leaves the automaton");
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->cg_next), n * sizeof(int32_t)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->cg_delta), n * sizeof(float)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->cg_resid), (size_t)n_states * sizeof(float)));
  K2B_CUDA(h, cudaMemcpy(h->cg_next, next, n * sizeof(int32_t), cudaMemcpyHostToDevice));
  K2B_CUDA(h, cudaMemcpy(h->cg_delta, delta, n * sizeof(float), cudaMemcpyHostToDevice));
  K2B_CUDA(h, cudaMemcpy(h->cg_resid, residual, (size_t)n_states * sizeof(float), cudaMemcpyHostToDevice));
  h->cg_states = n_states;
  return K2B_OK;
}

// ---- streaming modified_beam_search ---------------------------------------------------------------------------------
int32_t k2b_beam_pool_create(k2b_handle* h, int32_t max_streams, int32_t K, int32_t max_frames) {
  K2B_TRY(enter(h));
  if (max_streams <= 0 || max_frames <= 0) return fail(h, K2B_ERR_INVALID, "k2b_beam_pool_create: max_streams and max_frames must be positive");
  if (K < 1 || K > kMaxBeam || K > h->cfg.vocab_size) return fail(h, K2B_ERR_INVALID, "k2b_beam_pool_create: beam must be in 1..8 and <= vocab_size");
  return beam_pool_create(h, max_streams, K, max_frames);
}

int32_t k2b_beam_pool_reset(k2b_handle* h, int32_t slot, const int64_t* hyp) {
  K2B_TRY(enter(h));
  return beam_pool_reset(h, slot, hyp);
}

int32_t k2b_modified_beam_search_online_chunk_dev(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t Tc,
                                                  const int32_t* slots, int64_t* hyp_out, int64_t* tokens, int32_t* ts,
                                                  int32_t* n_out, float* score, int32_t cap) {
  K2B_TRY(enter(h));
  K2B_TRY(need_weights(h));
  h->lens_active = false;
  if (B < 0 || Tc < 0) return fail(h, K2B_ERR_INVALID, "k2b_modified_beam_search_online_chunk: negative B or Tc");
  if (B == 0) return K2B_OK;
  if (slots == nullptr || tokens == nullptr || ts == nullptr || n_out == nullptr || score == nullptr || (Tc > 0 && enc == nullptr) || cap < 1)
    return fail(h, K2B_ERR_INVALID, "k2b_modified_beam_search_online_chunk: NULL argument or cap < 1");
  K2B_TRY(beam_pool_begin(h, slots, B, Tc));
  const int K = beam_pool_K(h);
  const int mask3 = 1;                   // the online loops do not emit the literal id 1 either (ref OnlineRecognizer.cs:181)
  if (Tc > 0) {
    const size_t NK = (size_t)B * K, J = h->cfg.joiner_dim;
    K2B_TRY(ensure(h, h->ws_bp, sizeof(int32_t) * NK * Tc));
    int32_t* bp = static_cast<int32_t*>(h->ws_bp.p);
    if (h->cfg.precision != K2B_PREC_FP32 && cluster_path_supported(h, K)) {
      K2B_TRY(ensure_cluster_assets(h));
      const size_t n = (size_t)B * Tc;
      K2B_TRY(ensure(h, h->ws_encproj, sizeof(float) * n * J));
      float* encE = static_cast<float*>(h->ws_encproj.p);
      if (enc_is_raw) {
        if (h->cfg.encoder_dim <= 0 || h->enc_w == nullptr) return fail(h, K2B_ERR_STATE, "encoder_proj weights not loaded (E == 0)");
        if (encproj_tc_supported(h)) {
          K2B_TRY(encoder_proj_tc(h, enc, (int)n, encE, true));
        } else {
          GemmArgs g;
          g.M = (int)n; g.N = (int)J; g.K = h->cfg.encoder_dim;
          g.A = enc; g.W = h->enc_w; g.bias = h->enc_b; g.C = encE;
          K2B_TRY(launch_gemm_simt(h, PRO_PLAIN, EPI_EXP2X, g));
        }
      } else {
        K2B_TRY(exp2x_frames(h, enc, encE, n * J));
      }
      const size_t a4 = (NK * 4 + 255) & ~size_t(255), a8 = (NK * 8 + 255) & ~size_t(255), ab = ((size_t)B * 4 + 255) & ~size_t(255);
      K2B_TRY(ensure(h, h->ws_state, 2 * a4 + ab + 2 * a8));
      char* p = static_cast<char*>(h->ws_state.p);
      BeamStateView v;
      v.lp = reinterpret_cast<float*>(p); p += a4;
      v.len = reinterpret_cast<int32_t*>(p); p += a4;
      v.nlive = reinterpret_cast<int32_t*>(p); p += ab;
      v.ctx = reinterpret_cast<int32_t*>(p); p += a8;
      v.hash = reinterpret_cast<unsigned long long*>(p);
      v.cst = K > 1 ? cluster_cst(h, B, K) : nullptr;
      K2B_TRY(beam_pool_gather(h, B, v));
      K2B_TRY(beam_cluster_dev(h, encE, B, Tc, K, bp, v.lp, v.len, v.nlive, mask3, nullptr, nullptr, 0, Tc, 1, v.ctx, v.hash, true));
      K2B_TRY(beam_pool_scatter(h, B, Tc, v, bp));
    } else {
      const float* frames = nullptr;
      K2B_TRY(frames_for_search(h, enc, enc_is_raw, B, Tc, &frames));
      K2B_TRY(ensure(h, h->ws_state, 2 * beam_state_bytes(B, K)));
      K2B_TRY(beam_pool_gather(h, B, beam_state_view(h, B, K, 0)));
      K2B_TRY(beam_dev(h, frames, B, Tc, K, tokens, ts, n_out, score, cap, mask3, nullptr, false, 0, 0, true));
      K2B_TRY(beam_pool_scatter(h, B, Tc, beam_state_view(h, B, K, Tc & 1), bp));
    }
  }
  return beam_pool_backtrace(h, B, tokens, ts, n_out, score, hyp_out, cap);
}

int32_t k2b_modified_beam_search_online_chunk(k2b_handle* h, const float* enc, int32_t enc_is_raw, int32_t B, int32_t Tc,
                                              const int32_t* slots, int64_t* hyp_out, int64_t* tokens, int32_t* ts,
                                              int32_t* n_out, float* score, int32_t cap) {
  K2B_TRY(enter(h));
  StagedInputGuard staged_guard(h);
  K2B_TRY(need_weights(h));
  if (B < 0 || Tc < 0) return fail(h, K2B_ERR_INVALID, "k2b_modified_beam_search_online_chunk: negative B or Tc");
  if (B == 0) return K2B_OK;
  if (slots == nullptr || tokens == nullptr || ts == nullptr || n_out == nullptr || score == nullptr || (Tc > 0 && enc == nullptr) || cap < 1)
    return fail(h, K2B_ERR_INVALID, "k2b_modified_beam_search_online_chunk: NULL argument or cap < 1");
  const size_t width = enc_is_raw ? h->cfg.encoder_dim : h->cfg.joiner_dim;
  const size_t in_bytes = sizeof(float) * (size_t)B * Tc * width;
  K2B_TRY(ensure(h, h->ws_in, in_bytes > 0 ? in_bytes : 4));
  OutStage o;
  K2B_TRY(stage_out(h, B, cap, &o));
  if (in_bytes) K2B_TRY(h2d_rows(h, h->ws_in.p, in_bytes, enc, in_bytes, in_bytes, 1, h->stream));
  K2B_TRY(k2b_modified_beam_search_online_chunk_dev(h, static_cast<const float*>(h->ws_in.p), enc_is_raw, B, Tc, slots,
                                                    hyp_out ? o.hyp : nullptr, o.tokens, o.ts, o.n, o.score, cap));
  K2B_CUDA(h, cudaMemcpyAsync(tokens, o.tokens, sizeof(int64_t) * (size_t)B * cap, cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaMemcpyAsync(ts, o.ts, sizeof(int32_t) * (size_t)B * cap, cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaMemcpyAsync(n_out, o.n, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaMemcpyAsync(score, o.score, sizeof(float) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  if (hyp_out) K2B_CUDA(h, cudaMemcpyAsync(hyp_out, o.hyp, sizeof(int64_t) * 2 * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  return finish_host_call(h);
}

// diagnostic / parity evidence: the back-pointer rows of the LAST beam search of this handle ([B,T,K] int32, entry =
// (parent slot << 28) | (appended token + 1), 0 for a dead slot), i.e. the whole history of the beam
int32_t k2b_debug_backpointers(k2b_handle* h, int32_t* out, int32_t B, int32_t T, int32_t K) {
  K2B_TRY(enter(h));
  const size_t bytes = sizeof(int32_t) * (size_t)B * T * K;
  if (out == nullptr || B < 0 || T < 0 || K < 1 || bytes > h->ws_bp.bytes) return fail(h, K2B_ERR_INVALID, "k2b_debug_backpointers: no such history");
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  K2B_CUDA(h, cudaMemcpy(out, h->ws_bp.p, bytes, cudaMemcpyDeviceToHost));
  return K2B_OK;
}

int32_t k2b_ctc_greedy_dev(k2b_handle* h, const float* logp, int32_t B, int32_t T, int32_t V, int32_t blank,
                           const int32_t* frame_offset, int64_t* prev_inout, int64_t* tokens, int32_t* ts, int32_t* n_out,
                           int32_t* trailing_blank_inout, int32_t cap) {
  K2B_TRY(enter(h));
  K2B_TRY(check_search_args(h, logp, B, T, cap, tokens, ts, n_out, "k2b_ctc_greedy"));
  if (V < 1) return fail(h, K2B_ERR_INVALID, "k2b_ctc_greedy: V must be >= 1");
  if (B == 0) return K2B_OK;
  return ctc_greedy_dev(h, logp, B, T, V, blank, frame_offset, prev_inout, tokens, ts, n_out, trailing_blank_inout, cap);
}

int32_t k2b_ctc_greedy(k2b_handle* h, const float* logp, int32_t B, int32_t T, int32_t V, int32_t blank,
                       const int32_t* frame_offset, int64_t* prev_inout, int64_t* tokens, int32_t* ts, int32_t* n_out,
                       int32_t* trailing_blank_inout, int32_t cap) {
  K2B_TRY(enter(h));
  K2B_TRY(check_search_args(h, logp, B, T, cap, tokens, ts, n_out, "k2b_ctc_greedy"));
  if (V < 1) return fail(h, K2B_ERR_INVALID, "k2b_ctc_greedy: V must be >= 1");
  if (B == 0) return K2B_OK;
  const size_t in_bytes = sizeof(float) * (size_t)B * T * V;
  K2B_TRY(ensure(h, h->ws_in, in_bytes > 0 ? in_bytes : 4));
  OutStage o;
  K2B_TRY(stage_out(h, B, cap > 0 ? cap : 1, &o));
  if (in_bytes) K2B_TRY(h2d_rows(h, h->ws_in.p, in_bytes, logp, in_bytes, in_bytes, 1, h->stream));
  if (frame_offset) K2B_CUDA(h, cudaMemcpyAsync(o.aux_a, frame_offset, sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
  if (trailing_blank_inout) K2B_CUDA(h, cudaMemcpyAsync(o.aux_b, trailing_blank_inout, sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
  if (prev_inout) K2B_CUDA(h, cudaMemcpyAsync(o.prev, prev_inout, sizeof(int64_t) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
  K2B_TRY(k2b_ctc_greedy_dev(h, static_cast<const float*>(h->ws_in.p), B, T, V, blank, frame_offset ? o.aux_a : nullptr,
                             prev_inout ? o.prev : nullptr, o.tokens, o.ts, o.n, trailing_blank_inout ? o.aux_b : nullptr, cap));
  if (cap > 0) {
    K2B_CUDA(h, cudaMemcpyAsync(tokens, o.tokens, sizeof(int64_t) * (size_t)B * cap, cudaMemcpyDeviceToHost, h->stream));
    K2B_CUDA(h, cudaMemcpyAsync(ts, o.ts, sizeof(int32_t) * (size_t)B * cap, cudaMemcpyDeviceToHost, h->stream));
  }
  K2B_CUDA(h, cudaMemcpyAsync(n_out, o.n, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  if (trailing_blank_inout) K2B_CUDA(h, cudaMemcpyAsync(trailing_blank_inout, o.aux_b, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  if (prev_inout) K2B_CUDA(h, cudaMemcpyAsync(prev_inout, o.prev, sizeof(int64_t) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  if (h->opt_async_d2h) return K2B_OK;
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  return K2B_OK;
}

// ---- page-locked host memory ------------------------------------------------------------------------------------
// A P/Invoke float[] is pageable: its copies are staged through the driver's bounce buffer and serialise with the host. Callers
// that want the PCIe rate (and the async_d2h overlap) take their frame / result buffers from here, or register their own.
int32_t k2b_host_alloc(void** out, int64_t bytes) {
  if (out == nullptr || bytes <= 0) return K2B_ERR_INVALID;
  *out = nullptr;
  return cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable) == cudaSuccess ? K2B_OK : (cudaGetLastError(), K2B_ERR_CUDA);
}
int32_t k2b_host_free(void* p) {
  if (p == nullptr) return K2B_OK;
  return cudaFreeHost(p) == cudaSuccess ? K2B_OK : (cudaGetLastError(), K2B_ERR_CUDA);
}
int32_t k2b_host_register(void* p, int64_t bytes) {
  if (p == nullptr || bytes <= 0) return K2B_ERR_INVALID;
  return cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable) == cudaSuccess ? K2B_OK : (cudaGetLastError(), K2B_ERR_CUDA);
}
int32_t k2b_host_unregister(void* p) {
  if (p == nullptr) return K2B_OK;
  return cudaHostUnregister(p) == cudaSuccess ? K2B_OK : (cudaGetLastError(), K2B_ERR_CUDA);
}

}  // extern "C"
