// Internal definitions shared by the translation units of libk2b200.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/k2b200.h"

namespace k2b {

constexpr int kMaxBeam = 8;

// SIMT GEMM tile (gemm_simt.cu); the partial-result layouts of the search kernels depend on BN.
constexpr int kBM = 64;
constexpr int kBN = 64;
constexpr int kBK = 16;

// A device allocation that only ever grows; freed with the handle.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct ProfEvents {
  std::vector<cudaEvent_t> start, stop;
  size_t used = 0;
  int64_t n_total = 0;
  double ms_total = 0.0;
};

}  // namespace k2b

namespace k2b { struct StatePool; struct BeamPool; struct NcclState; struct HostStage; }

struct k2b_handle {
  k2b_config cfg{};
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  bool poisoned = false;
  bool weights_loaded = false;
  int64_t launches = 0;
  int sm_count = 148;

  // fp32 weights as loaded (row-major, layouts in k2b200.h)
  float* emb = nullptr;      // [V,D]
  float* conv_w = nullptr;   // [D,4,ctx]
  float* dec_w = nullptr;    // [J,D]
  float* dec_b = nullptr;    // [J]
  float* enc_w = nullptr;    // [J,E]
  float* enc_b = nullptr;    // [J]
  float* out_w = nullptr;    // [V,J]
  float* out_b = nullptr;    // [V]
  // grouped conv folded into per-token tables at load time: conv(e0,e1)[o] = T0[y0][o] + T1[y1][o].
  // Row V of each table is all-zero (the "masked" negative id).
  float* tab0 = nullptr;     // [V+1,D]
  float* tab1 = nullptr;     // [V+1,D]

  // growable workspaces
  k2b::DevBuf ws_in;        // host-variant staging of the input frames / log-probs
  k2b::DevBuf ws_encproj;   // [B,T,J] projected frames when the caller passes raw frames
  k2b::DevBuf ws_x;         // [N,J] joiner A operand tanh(enc+dec)
  k2b::DevBuf ws_ximg;      // the same operand as bf16 hi/lo tile images (written by the tcgen05 decoder, read by TMA)
  k2b::DevBuf ws_dec;       // [N,J] decoder_proj output (fine-grained path)
  k2b::DevBuf ws_logits;    // [N,V] (fine-grained path staging)
  k2b::DevBuf ws_part;      // per-(row, vocab tile) partial argmax / softmax / top-k
  k2b::DevBuf ws_state;     // hypothesis state (ctx, lp, len, hash, nlive), double-buffered
  k2b::DevBuf ws_bp;        // beam back-pointers [B,T,K]
  k2b::DevBuf ws_out;       // host-variant staging of tokens / ts / n / score
  k2b::DevBuf ws_misc;      // int32 contexts of the fine-grained decoder call
  k2b::DevBuf ws_sync;      // persistent beam-search kernel: per-(frame, row tile) counters
  k2b::DevBuf ws_ctc;       // ctc: per-frame ids [B*T] + per-stream tickets [B] (tickets stay zero between launches)

  // tensor-core (cluster) path assets, built on first use after a weight load (search_cluster.cu)
  bool tc_ready = false;
  uint8_t* wo_hi_img = nullptr;   // [CS][J/64][16 KB] bf16 hi of out_w in the swizzled shared-memory image
  uint32_t* wo_lo = nullptr;      // [CS*128][J/2] bf16 lo of out_w, packed pairs (TMEM source)
  uint32_t* wo_hi_rows = nullptr; // [CS*128][J/2] bf16 hi of out_w, packed pairs (TMEM source of the k-blocks kept in tensor memory)
  float* bias_pad = nullptr;      // [CS*128] out_b, -inf beyond V
  float* dec_tab = nullptr;       // [(V+1)*V, J] exp(2*decoder(y0,y1)): the memoised stateless decoder
  int dec_tab_state = 0;          // 0 not tried, 1 built, -1 does not fit (decoder GEMM per frame instead)
  size_t dec_tab_bytes = 0;       // what the memoised decoder occupies, and the device time its build took (k2b_get_stat)
  float dec_tab_build_ms = 0.f;

  int cluster16_ok = -1;          // 16-CTA cluster launchable on this device? (-1 unknown; occupancy query, cached)
  bool enc_ready = false;         // encproj_tc.cu: pre-split, pre-swizzled encoder_proj weight images
  uint8_t* we_hi_img = nullptr;
  uint8_t* we_lo_img = nullptr;
  bool wj_ready = false;          // out_w in 256-row tiles for the per-frame tcgen05 joiner (any vocabulary)
  uint8_t* wj_hi_img = nullptr;
  uint8_t* wj_lo_img = nullptr;
  bool wd_ready = false;          // dec_proj_w in 256-row tiles for the tcgen05 decoder (per-frame path)
  uint8_t* wd_hi_img = nullptr;
  uint8_t* wd_lo_img = nullptr;
  cudaStream_t copy_stream = nullptr;   // host-pointer calls: H2D of time chunk c+1 overlaps compute of chunk c
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  // "inputs_complete" look-ahead of the device-pointer beam search (api.cu::beam_cluster_lookahead): both encoder_proj GEMMs of a
  // call run on a low-priority side stream, i.e. under the PREVIOUS call's search; two projected-frame buffers taken in turns
  cudaStream_t la_stream = nullptr;
  cudaEvent_t la_ev_a = nullptr, la_ev_b = nullptr, la_ev_done[2] = {nullptr, nullptr};
  bool la_done_valid[2] = {false, false};
  unsigned la_turn = 0;
  k2b::DevBuf ws_encproj_la[2];
  // recorded on the handle's stream right in front of the search kernel of a look-ahead call: the NEXT call's side-stream work waits
  // for it, so that its low-priority CTAs do not take the SMs in the gap before that search is launched (they would have to
  // drain first: up to one GEMM tile, ~20 us, per launch)
  cudaEvent_t la_ev_mark = nullptr;
  bool la_mark_valid = false, la_mark_arm = false;
  int mega_grid_cap = 0;                  // > 0: the persistent beam kernel of the current call takes at most this many SMs
  int* la_flags = nullptr;                // [kLaFlags] "time chunk c is projected" epochs, polled by the running cluster kernel
  int la_epoch = 0;
  int* dev_status = nullptr;      // [4] error flags written by the tcgen05 kernels (mbarrier time-outs)
  long long* cluster_timing = nullptr;   // device [8]: per-phase cycle totals of the cluster kernel (diagnostic)

  long long* timeline = nullptr;  // diagnostic: [64 frames][148 SMs][8] clock64 stamps of the per-frame beam path (k2b_debug_timeline)
  int timeline_frame = 0;
  bool profile_on = false;
  int prof_which = 0;             // which launch of the per-frame beam path the profile events bracket (K2B_PROF_WHICH: 0 joiner, 1 operand build, 2 merge)
  k2b::StatePool* state_pool = nullptr;   // on-device streaming state (state_pool.cu)
  int32_t* lens_dev = nullptr;    // k2b_set_encoder_out_lens: per-stream frame counts for the next fused offline search
  int lens_n = 0;
  bool lens_active = false;
  k2b::ProfEvents prof;
  k2b::BeamPool* beam_pool = nullptr;     // streaming modified_beam_search: per-stream hypotheses carried between chunks (stream_beam.cu)
  k2b::NcclState* nccl = nullptr;         // k2b_nccl_init / k2b_gather_results_nccl (nccl_gather.cu; libnccl is dlopen'ed)
  // contextual biasing (hot words) of modified_beam_search: a dense automaton over token ids (k2b_set_context_graph)
  int32_t* cg_next = nullptr;             // [S,V] next state
  float* cg_delta = nullptr;              // [S,V] score added on that transition
  float* cg_resid = nullptr;              // [S]   unearned partial boost revoked when the utterance ends in state s
  int cg_states = 0;
  k2b::DevBuf ws_cst;                     // cluster engine: automaton state per hypothesis [B*K], carried between time chunks
  k2b::HostStage* host_stage = nullptr;   // page-locked bounce buffers + copy threads for pageable inputs (host_stage.cu)
  int max_sym_per_frame = 1;              // k2b_set_option("max_sym_per_frame"): ref OfflineRecognizer.cs:19 fixes it to 1
  // engine switches (k2b_set_option; the K2B_* environment variables only give their initial values at k2b_create)
  int opt_la_frames = 0;                  // frames per time chunk of the polling device-pointer search (0 = chosen per call)
  int opt_inputs_complete = 0;            // 1: frames handed to _dev calls are complete in memory at call time (not merely stream-ordered)
  int opt_async_gather = 0;               // 1: k2b_gather_results_nccl runs on a side stream (k2b_gather_join / k2b_sync order behind it)
  int opt_copy_threads = -1;              // host threads staging pageable inputs (-1: a quarter of the host's, 2 .. 8)
  int opt_pipe_chunks = 0;                // > 0: number of time chunks of the host-pointer beam search
  int opt_no_mega = 0;                    // large-vocabulary beam search as per-frame launches instead of the persistent launch
  int opt_unfused_step = 0;               // the three-launch frame step
  int opt_greedy_persistent = -1;         // -1 auto, 0 cluster kernel, 1 persistent kernel (greedy, 1024 < V <= 2048)
  int opt_pair = 0;                       // CTA-pair variant of the cluster kernel
  int opt_ctc_one_kernel = -1;            // CTC greedy: 1 = one kernel (tickets), 0 = frames + collapse kernels (PDL), -1 = by input size
  // persistent greedy kernel: records carry the frame's epoch tag instead of being announced by a counter (joiner_tc.cu); the tags
  // of a launch are ll_epoch + 1 .. ll_epoch + T; ll_clean_ptr / _bytes: the partials buffer as it was last zeroed for tagged use
  uint32_t ll_epoch = 1;
  void* ll_clean_ptr = nullptr;
  size_t ll_clean_bytes = 0;
  int ll_clean_kk = 0;                    // ... and the record layout (beam bound of the instantiation) it was zeroed for
  int opt_tagged_records = 1;             // persistent kernels: 0 = records announced through counters, 1 = epoch-tagged records
  int opt_single_greedy = 1;              // greedy search of up to 16 streams on the register-resident fp32 kernel (single_greedy.cu)
  int opt_dev_chunks = -1;                // device-pointer beam search on the cluster kernel: time chunks (-1 = two when it pays, 1 = one)
  int opt_wh_tmem = -1;                   // cluster kernel: k-blocks of W_hi held in tensor memory (-1 = balanced choice)
  int opt_async_d2h = 0;                  // host-pointer fused calls return without the final sync (pinned buffers; k2b_sync completes)
};

namespace k2b {

int32_t fail(k2b_handle* h, int32_t code, const std::string& msg);
int32_t cuda_fail(k2b_handle* h, cudaError_t e, const char* what, const char* file, int line);
int32_t ensure(k2b_handle* h, DevBuf& b, size_t bytes);

// Launch with programmatic stream serialization: the kernel may begin while the previous kernel of the stream drains (it calls
// griddep_wait() before reading that kernel's output). Used for the back-to-back launches of the per-frame search path.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define K2B_CUDA(h, expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return k2b::cuda_fail((h), _e, #expr, __FILE__, __LINE__); \
  } while (0)

#define K2B_TRY(expr)                   \
  do {                                  \
    int32_t _s = (expr);                \
    if (_s != K2B_OK) return _s;        \
  } while (0)

#define K2B_LAUNCH_CHECK(h)                                                      \
  do {                                                                           \
    (h)->launches++;                                                             \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) return k2b::cuda_fail((h), _e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

// ---- gemm_simt.cu ---------------------------------------------------------------------------
enum Pro : int { PRO_PLAIN = 0, PRO_DEC = 1, PRO_JOIN = 2 };
enum Epi : int { EPI_STORE = 0, EPI_TANH_ADD = 1, EPI_ARGMAX = 2, EPI_TOPK = 3, EPI_EXP2X = 4 };

struct GemmArgs {
  int M = 0, N = 0, K = 0;
  const float* W = nullptr;     // [N,K]
  const float* bias = nullptr;  // [N] or null
  // PRO_PLAIN
  const float* A = nullptr;     // [M,K]
  // PRO_DEC: A(m,k) = relu(tab0[r0(m)][k] + tab1[r1(m)][k]), ids from ctx[m][2]
  const int32_t* ctx = nullptr;
  const float* tab0 = nullptr;
  const float* tab1 = nullptr;
  int V = 0;
  int neg_wrap = 0;
  int blank = 0;
  const int32_t* compat_flag = nullptr;  // Q6: when *flag != 0 a negative ctx[m][0] reads as blank
  // PRO_JOIN / EPI_TANH_ADD: encoder row of GEMM row m is enc + (m / rows_per_stream) * enc_stride
  const float* enc = nullptr;
  long long enc_stride = 0;
  int rows_per_stream = 1;
  const float* dec = nullptr;   // PRO_JOIN: [M,K]
  // EPI_STORE / EPI_TANH_ADD
  float* C = nullptr;           // [M,N]
  // EPI_ARGMAX: [M,nt] each
  float* part_val = nullptr;
  int32_t* part_idx = nullptr;
  int32_t* part_nan = nullptr;
  // EPI_TOPK: part_m/part_s [M,nt]; part_tv/part_ti [M,nt,topk]
  float* part_m = nullptr;
  float* part_s = nullptr;
  float* part_tv = nullptr;
  int32_t* part_ti = nullptr;
  int topk = 0;
};

int32_t launch_gemm_simt(k2b_handle* h, Pro pro, Epi epi, const GemmArgs& a);
inline int num_vocab_tiles(int V) { return (V + kBN - 1) / kBN; }

// ---- decoder_tables.cu -----------------------------------------------------------------------
int32_t build_decoder_tables(k2b_handle* h);

// ---- ctc.cu ------------------------------------------------------------------------------------
int32_t ctc_greedy_dev(k2b_handle* h, const float* logp, int B, int T, int V, int blank,
                       const int32_t* frame_offset, int64_t* prev_inout, int64_t* tokens, int32_t* ts,
                       int32_t* n_out, int32_t* trailing_inout, int cap);

// ---- search.cu ---------------------------------------------------------------------------------
int32_t greedy_dev(k2b_handle* h, const float* enc, int B, int T, int mode, bool online,
                   int64_t* hyp_inout, int64_t* tokens, int32_t* ts, int32_t* n_out, int cap);
// extra_mask / hyp_inout: greedy search as beam 1 (the literal-1 mask and OnlineStream.Hyp of the online loop); only the engines
// built on beam_merge_stream implement them - beam_dev fails with K2B_ERR_UNSUPPORTED otherwise (beam_greedy_usable tells)
// carry: the hypothesis state is already in the first state buffer (beam_state_ptrs(.., 0)); T frames are decoded from it, the final
// state is left in buffer T & 1 and NO back-trace runs (streaming beam search: the caller owns history and state).
// enc_stride: floats between the frames of consecutive streams when it is not Ttot * J.
int32_t beam_dev(k2b_handle* h, const float* enc, int B, int T, int K, int64_t* tokens, int32_t* ts,
                 int32_t* n_out, float* score, int cap, int extra_mask = -1, int64_t* hyp_inout = nullptr, bool greedy = false,
                 int t0 = 0, int Ttot = 0, bool carry = false, long long enc_stride = 0);
struct BeamStateView { int32_t* ctx; float* lp; int32_t* len; unsigned long long* hash; int32_t* nlive; int32_t* cst = nullptr; };
size_t beam_state_bytes(int B, int K);                                   // one of the two state buffers of beam_dev
BeamStateView beam_state_view(k2b_handle* h, int B, int K, int which);   // inside ws_state (ensure 2 * beam_state_bytes first)
// a search may be stepped in time chunks (enc = frame t0 of a [B,Ttot,J] array, T frames per call: the host-pointer entry point hides
// the input copy behind the search that way) on the engines built on beam_merge_stream
bool beam_chunkable(k2b_handle* h, int K);
bool beam_greedy_usable(k2b_handle* h);

int32_t beam_backtrace_dev(k2b_handle* h, int B, int K, int T, const float* lp, const int32_t* len, const int32_t* nlive,
                           const int32_t* bp, int64_t* tokens, int32_t* ts, int32_t* n_out, float* score, int cap,
                           const int32_t* cst = nullptr);
int32_t* cluster_cst(k2b_handle* h, int B, int K);      // the cluster engine's automaton-state array (null without a context graph)

// ---- single_greedy.cu --------------------------------------------------------------------------
bool single_greedy_usable(const k2b_handle* h, int B, int T);
int32_t single_greedy_dev(k2b_handle* h, const float* encE, int B, int T, int t0, int Ttot, int extra_mask, const int64_t* hyp_in,
                          int64_t* hyp_out, int64_t* tokens, int32_t* ts, int32_t* n_out, int cap);

// ---- search_cluster.cu -------------------------------------------------------------------------
bool cluster_path_supported(const k2b_handle* h, int K);
int32_t ensure_cluster_assets(k2b_handle* h);
int32_t ensure_dec_table(k2b_handle* h, bool* have);
int32_t joinin_table_tc(k2b_handle* h, const int32_t* ctx, int M, const float* enc, long long enc_stride, int rows_per_stream,
                        uint8_t* x_img);
int32_t exp2x_frames(k2b_handle* h, const float* in, float* out, size_t n);
int32_t exp2x_frames_chunk(k2b_handle* h, const float* in, float* out, int B, int tc, int T, int t0);   // [B,tc,J] -> rows t0.. of [B,T,J]
int32_t beam_cluster_dev(k2b_handle* h, const float* encE, int B, int T, int K, int32_t* bp, float* fin_lp, int32_t* fin_len,
                         int32_t* fin_nlive, int extra_mask, const int64_t* hyp_in, int64_t* hyp_out, int t0 = 0, int Ttot = 0,
                         int resume = 0, int32_t* io_ctx = nullptr, unsigned long long* io_hash = nullptr, bool need_lp = true,
                         const int* ready = nullptr, int ready_epoch = 0, int ready_len = 0);
int cluster_grid_ctas(const k2b_handle* h, int B, int K);            // CTAs of one cluster-kernel launch (all resident at once or not)
int32_t cluster_set_ready(k2b_handle* h, int* flag, int epoch);      // on h->stream: *flag = epoch, released at device scope
int32_t cluster_status(k2b_handle* h);

// ---- stream_beam.cu: hypotheses carried between the chunks of a stream -----------------------------------------
int32_t beam_pool_create(k2b_handle* h, int max_streams, int K, int max_frames);
void beam_pool_free(k2b_handle* h);
int32_t beam_pool_reset(k2b_handle* h, int slot, const int64_t* hyp_host);
// slots_host [B]; layout A (cluster kernel): separate arrays; layout B (beam_dev): a BeamStateView
int32_t beam_pool_begin(k2b_handle* h, const int32_t* slots_host, int B, int Tc);
int32_t beam_pool_gather(k2b_handle* h, int B, const BeamStateView& dst);
int32_t beam_pool_scatter(k2b_handle* h, int B, int Tc, const BeamStateView& src, const int32_t* bp_chunk);
int32_t beam_pool_backtrace(k2b_handle* h, int B, int64_t* tokens, int32_t* ts, int32_t* n_out, float* score, int64_t* hyp_out, int cap);
int beam_pool_K(const k2b_handle* h);

// ---- host_stage.cu: H2D of `rows` rows of `width` bytes; pageable sources are staged by the library's own copy threads ----------
int32_t h2d_rows(k2b_handle* h, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaStream_t stream);
void host_stage_free(k2b_handle* h);

// ---- nccl_gather.cu ------------------------------------------------------------------------------------------
void nccl_free(k2b_handle* h);
int32_t gather_join(k2b_handle* h);       // the handle's stream waits for an outstanding side-stream gather ("async_gather")

// ---- encproj_tc.cu -----------------------------------------------------------------------------
bool encproj_tc_supported(const k2b_handle* h);
int32_t ensure_encproj_assets(k2b_handle* h);
// rows_per_stream > 0: the n = B * rows_per_stream rows are time chunk [out_t0, out_t0 + rows_per_stream) of every stream and go to
// rows b * out_T + out_t0 + tt of `out`; in_T > 0: `raw` is the whole [B,in_T,E] array as well (otherwise a compact [B,chunk,E] one)
int32_t encoder_proj_tc(k2b_handle* h, const float* raw, int n, float* out, bool exp2x, int rows_per_stream = 0, int out_T = 0,
                        int out_t0 = 0, int in_T = 0);
int32_t state_pool_create(k2b_handle* h, const int32_t* item_len, int n_tensors, int max_streams);
void state_pool_free(k2b_handle* h);
int32_t state_pool_restack(k2b_handle* h, const int32_t* slots, int B, const int32_t* axis_len, float* stacked_dev, bool unstack);
int32_t state_pool_io(k2b_handle* h, int slot, float* host, bool put);
size_t state_pool_stacked_floats(const k2b_handle* h, int B);
bool decoder_tc_supported(const k2b_handle* h);
int32_t decoder_joinin_tc(k2b_handle* h, const int32_t* ctx, int M, const float* enc, long long enc_stride, int rows_per_stream,
                          float* x, uint8_t* x_img);
size_t joiner_tc_image_bytes(const k2b_handle* h, int M);
bool joiner_tc_supported(const k2b_handle* h);
int32_t ensure_joiner_assets(k2b_handle* h);
int joiner_tc_tiles(const k2b_handle* h);
int32_t joiner_tc_partials(k2b_handle* h, const float* x, const uint8_t* x_img, int M, int topk, float* part_m, float* part_s,
                           float* part_tv, int32_t* part_ti, float* part_val, int32_t* part_idx, int32_t* part_nan);

// ---- joiner_tc.cu ------------------------------------------------------------------------------
bool joiner_topk_supported(const k2b_handle* h, int topk);
bool joiner_topk_usable(const k2b_handle* h, int topk);   // the persistent joiner will serve joiner_tc_partials(x_img, topk)
int32_t joiner_topk_tc(k2b_handle* h, const uint8_t* x_img, int M, int topk, int kk, float* part_rec);
int joiner_topk_tiles(const k2b_handle* h, int M);   // vocabulary tiles (= records per row) joiner_topk_tc / beam_mega_tc use for M rows
int beam_partial_words(int kk);        // floats per (row, tile) record of joiner_topk_tc / beam_mega_tc (kk = 1: greedy, 4, 8)

// beam 1 on the persistent kernel leaves its results itself (no back-pointers, no back-trace launch)
struct GreedyOutPtrs { int64_t* tokens; int32_t* ts; int32_t* n; int64_t* hyp; int cap; };
struct BeamStatePtrs { int32_t* ctx; float* lp; int32_t* len; unsigned long long* hash; int32_t* nlive; };
constexpr int32_t kMegaUnavailable = 0x4d454741;   // beam_mega_tc: the cooperative launch does not fit; nothing was enqueued
bool beam_mega_usable(const k2b_handle* h, int K);
int32_t beam_mega_tc(k2b_handle* h, const float* enc, int B, int T, int K, uint8_t* x_img, float* part_rec, const BeamStatePtrs& s0,
                     const BeamStatePtrs& s1, int32_t* bp, const int32_t* lens, int mask3, int kk, const GreedyOutPtrs* go, bool sync_zeroed,
                     int t0, int Ttot, long long enc_stride);
size_t beam_mega_sync_ints(const k2b_handle* h, int B, int T, int K);

// profiling bracket around the dominant GEMM
void prof_begin(k2b_handle* h);
void prof_end(k2b_handle* h);

}  // namespace k2b
