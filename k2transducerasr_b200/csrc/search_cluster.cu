// Persistent thread-block-cluster kernel: the whole modified_beam_search time loop of a group of streams runs
// inside ONE launch; nothing but the encoder frames (read) and the back-pointers (written) touches HBM per step.
//
// Mapping (V <= 1024, J <= 512):
//   * a cluster of CS = ceil(V/128) CTAs owns S = 32/K streams = NH = 32 hypothesis rows for all T frames;
//   * CTA `rank` keeps its 128-row slice of the joiner output weight resident for the whole launch:
//       hi bf16 part  -> shared memory, staged once by bulk TMA copies in the K-major SWIZZLE_128B image,
//       lo bf16 part  -> TENSOR MEMORY (256 of the 512 TMEM columns), used as the TMEM-resident A operand,
//     so the fp32-accurate split-bf16 GEMM (Wh*xh + Wh*xl + Wl*xh) needs no second shared-memory copy;
//   * operands are swapped - D[vocab 128, hyps 32] = W_slice[128,J] * x[32,J]^T - so the vocabulary fills the
//     M = 128 lanes of tcgen05.mma and the few hypothesis rows are the (cheap) N dimension;
//   * the stateless decoder is memoised at weight-load time: dec_tab[(y0+1)*V + y1] = exp(2*decoder(y0,y1))
//     for every context (513 MB at V=500 - HBM capacity traded for a GEMM per step), the frames arrive as
//     exp(2*enc), and the joiner prologue tanh(enc+dec) = 1 - 2/(1 + Ee*Ed) costs one MUFU per element;
//   * epilogue: TMEM -> registers (+bias) -> shared transpose (all 16 worker warps) -> per hypothesis sum-exp (fixed-point REDUX)
//     and top-K (unique packed keys, one REDUX per round); the (2+2K)-word record of every (slice, hypothesis) goes to every
//     other CTA of the cluster with st.async stores that complete on the destination's mbarrier - no barrier.cluster in the
//     frame loop; every CTA then runs the hypothesis merge (log_softmax constants, per-stream top-K, dedupe by sequence hash,
//     log-add, prune) redundantly and deterministically, so no second exchange is needed. CTA 0 records the back-pointers.
//   * 17 warps: 16 workers + one that only issues the MMAs, K-quarter by K-quarter as the operand lands.
#include <math.h>
#include <stdlib.h>

#include "k2b_internal.h"
#include "sm100_ptx.cuh"

namespace k2b {

namespace {

using namespace ptx;

#ifndef K2B_CLUSTER_MAXREG
#define K2B_CLUSTER_MAXREG 96     // 17 warps: up to 120 would fit the register file
#endif
constexpr int kNH = 32;          // hypothesis rows per cluster (UMMA N)
constexpr int kWorkers = 16;     // warps 0..15: prologue / read-out / reductions / merge
constexpr int kCAll = (kWorkers + 1) * 32;   // + warp 16, which only issues MMAs
constexpr int kLtStride = 33;
constexpr int kBpSteps = 32;     // back-pointer rows are collected in shared memory and written out every kBpSteps frames
constexpr uint64_t kHashSeedC = 0x9E3779B97F4A7C15ull;

struct HypState {
  int ctx0[kNH], ctx1[kNH];
  float lp[kNH];
  int len[kNH];
  unsigned long long hash[kNH];
  int nlive[kNH];
  int cst[kNH];             // context-graph (hot word) state of every hypothesis; all zero without a graph
};

struct ClusterArgs {
  const float* encE;        // [B,T,J]  exp(2*clamp(enc))
  const float* dec_tab;     // [(V+1)*V, J]  exp(2*clamp(decoder(y0,y1)))
  const uint8_t* wo_hi_img; // [CS][J/64][128 x 128 B swizzled]
  const uint32_t* wo_lo;    // [CS*128][J/2] packed bf16 pairs
  const uint32_t* wo_hi_rows;  // [CS*128][J/2] packed bf16 pairs (source of the W_hi k-blocks kept in tensor memory)
  int nt;                   // the first nt k-blocks (64 joiner columns each) of W_hi live in TMEM, the others in shared memory
  const float* bias;        // [CS*128], -inf beyond V
  int B, T, K, V, J, S, CS, blank, unk, x3;
  int extra_mask;           // third non-emitting id (the literal 1 of ref OnlineRecognizer.cs:181), or -1
  int need_lp;              // 0: greedy search (beam 1, no score wanted): log-softmax is monotone, so raw logits are ranked
  const int64_t* hyp_in;    // [B,2] initial contexts (online: OnlineStream.Hyp), or null = {-1, blank}
  int64_t* hyp_out;         // [B,2] final context of slot 0 (online), or null
  // time-chunked operation: this launch decodes frames [t0, t0+T) of utterances Ttot frames long; with resume != 0 the
  // hypothesis state is read from fin_lp / fin_len / fin_nlive / io_ctx / io_hash (written by the previous launch)
  int t0, Ttot, resume;
  int32_t* io_ctx;          // [B*K,2]
  unsigned long long* io_hash;  // [B*K]
  int32_t* bp;              // [B,T,K]
  float* fin_lp;            // [B*K]
  int32_t* fin_len;         // [B*K]
  int32_t* fin_nlive;       // [B]
  int* status;
  long long* timing;        // optional [8] cycle totals of the step phases (cluster 0, CTA 0, thread 0)
  const int32_t* lens;      // [B] frames to decode per stream (k2b_set_encoder_out_lens), or null = all T
  // contextual biasing (hot words): dense automaton over token ids, or null. io_cst [B*K]: state per hypothesis, written at the end
  // of every launch and read back when `resume` is set
  const int32_t* cg_next;   // [S,V]
  const float* cg_delta;    // [S,V]
  int32_t* io_cst;
  // frames that arrive while the kernel runs (the encoder_proj GEMMs of later time chunks are on a side stream): frames
  // [c * ready_len, (c + 1) * ready_len) may be read once ready[c] has reached ready_epoch; null = all frames are there.
  // The chunk of frames t0 and t0 + 1 must be complete before the launch. The MMA warp holds back the commit of step t until the
  // chunk of frame t + 2 is there: the worker warps load frame t + 2 (one step ahead of its use) only behind that commit, so
  // their code is the same with and without flags (a wait in the workers' own loop cost 80 us per 250-frame launch - the
  // compiler split the loop -, whatever the kind of load).
  const int* ready;
  int ready_epoch, ready_len;
};

// polled by the MMA warp (all lanes, one address), bounded like the mbarrier waits
__device__ __forceinline__ bool wait_frames_ready(const int* flag, int epoch) {
  const long long t0 = clock64();
#pragma unroll 1
  for (;;) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v - epoch >= 0) return true;
    if (clock64() - t0 > kWaitTimeoutCycles) return false;
    __nanosleep(100);
  }
}

__device__ __forceinline__ bool better_c(float v, int i, float ev, int ei) { return v > ev || (v == ev && i > ei); }

__host__ __device__ constexpr int xw_padded(int K) { return ((2 + 2 * K + 3) / 4) * 4; }   // words per partial, 16 B multiple
// Bytes of the joiner-operand region: nkb k-block tiles, but never less than the transposed logits tile [128][33] fp32 that
// aliases it between the MMA and the next prologue (matters for J < 192).
__host__ __device__ constexpr size_t xop_region_bytes(int nkb, int tile_bytes, int lt_stride) {
  const size_t ops = (size_t)nkb * tile_bytes, lt = ((size_t)128 * lt_stride * 4 + 1023) / 1024 * 1024;
  return ops > lt ? ops : lt;
}

__device__ __forceinline__ uint64_t hash_push_c(uint64_t h, int tok) {
  h = (h ^ (uint64_t)(uint32_t)(tok + 1)) * 0x100000001B3ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h;
}
__device__ __forceinline__ float logaddexp_c(float a, float b) {
  const float mx = fmaxf(a, b), mn = fminf(a, b);
  if (mx == -INFINITY) return -INFINITY;
  return mx + __logf(1.f + __expf(mn - mx));       // |error| < 1e-7 absolute: the argument of log is in [1, 2]
}

// tanh(e + d) from Ee = exp(2e), Ed = exp(2d). e and d are clamped to +-21 when the tables are built, so the
// product stays finite (e^84 < FLT_MAX) and tanh is already saturated to 1 ulp far inside that range.
__device__ __forceinline__ float tanh_from_exp(float ee, float ed) {
  const float y = fmaf(ee, ed, 1.0f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ int fkey(float f) { const int k = __float_as_int(f); return k >= 0 ? k : (k ^ 0x7fffffff); }
__device__ __forceinline__ float funkey(int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff)); }
constexpr int kKeyNone = (int)0x80000000;

// One warp: hypothesis merge of local stream s (beam_select_kernel in search.cu is the global-memory twin).
// Lane c*K + h owns the record of (vocabulary slice c, hypothesis h) - its slice max / sum-exp and its K best (value, index)
// candidates - so the log-softmax constants are two unconditional butterflies over the slice bits of the lane id (lanes
// without a record contribute the neutral element) and every candidate is scored in the lane that loaded it. K rounds of
// two REDUX then pick the stream's top K over K*V (value, then flat index). The K winners meet in a small per-warp scratch
// (broadcast loads instead of shuffles) for the dedupe by token sequence (hash, length, context) and the log-add.
// CG: a context graph (hot words) is set - the only instantiation that touches the automaton state of the hypotheses.
template <int K, bool CG>
__device__ __forceinline__ void select_stream(int s, int V, int CS, const float* __restrict__ xb, const HypState& in,
                                              HypState& out, int blank, int unk, int extra_mask, int32_t* __restrict__ bp_row,
                                              int lane, const float* __restrict__ dec_tab, int J, bool do_prefetch,
                                              uint32_t* __restrict__ scr, bool need_lp, const int32_t* __restrict__ cg_next,
                                              const float* __restrict__ cg_delta, long long* tp = nullptr) {
#define K2B_SUB(i) do { if (tp != nullptr) { const long long now = clock64(); tp[i] += now - tp[19]; tp[19] = now; } } while (0)
  if (tp != nullptr) tp[19] = clock64();
  constexpr int XWP = xw_padded(K);
  constexpr int CPP = 32 / K;                     // slices per pass
  constexpr int NP = (K == 8) ? 2 : 1;            // K * CS <= 32 * NP on every supported shape
  const unsigned full = 0xffffffffu;
  const int nl = in.nlive[s];
  const int h = lane % K;
  float w[NP][XWP];
  float pm[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int c = p * CPP + lane / K;
    const float4* rec = reinterpret_cast<const float4*>(xb + ((size_t)min(c, CS - 1) * kNH + s * K + h) * XWP);
#pragma unroll
    for (int i = 0; i < XWP / 4; ++i) {
      const float4 q = rec[i];
      w[p][4 * i] = q.x; w[p][4 * i + 1] = q.y; w[p][4 * i + 2] = q.z; w[p][4 * i + 3] = q.w;
    }
  }
  const float lp_h = in.lp[s * K + h];
  if (nl == 0) {                                   // warp-uniform
    if (lane < K) {
      const int o = s * K + lane;
      out.ctx0[o] = -1; out.ctx1[o] = blank; out.lp[o] = -INFINITY; out.len[o] = 2; out.hash[o] = kHashSeedC;
      if (CG) out.cst[o] = 0;
    }
    if (lane == 0) out.nlive[s] = 0;
    if (bp_row != nullptr && lane < K) bp_row[lane] = 0;
    return;
  }
  K2B_SUB(8);
  float M = -INFINITY;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const bool valid = h < nl && p * CPP + lane / K < CS;
    pm[p] = valid ? w[p][0] : -INFINITY;
    M = fmaxf(M, pm[p]);
  }
  float L = 0.f;
  if (K > 1 || need_lp) {              // warp-uniform
#pragma unroll
    for (int o = K; o < 32; o <<= 1) M = fmaxf(M, __shfl_xor_sync(full, M, o));
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) sum += (pm[p] > -INFINITY) ? w[p][1] * __expf(pm[p] - M) : 0.f;
#pragma unroll
    for (int o = K; o < 32; o <<= 1) sum += __shfl_xor_sync(full, sum, o);
    L = __logf(sum);
  } else {
    M = 0.f;                           // greedy: rank the raw logits (one hypothesis, log-softmax is monotone)
  }
  int ck[NP * K], cf[NP * K];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int idx = __float_as_int(w[p][2 + K + j]);
      const float v = ((w[p][2 + j] - M) - L) + lp_h;                       // order of log_softmax(x) + lp
      const bool okc = (pm[p] > -INFINITY) & (idx >= 0) & (v == v);
      ck[p * K + j] = okc ? fkey(v) : kKeyNone;
      cf[p * K + j] = okc ? h * V + idx : -1;
    }
  }
  K2B_SUB(9);
  float my_v = -INFINITY;
  int my_f = -1;
  if constexpr (NP * K <= 4) {
    // few candidates per lane: sort them once (value, then flat index, best first); a round is then two REDUX and a pop
#define K2B_CE2(x, y)                                                                      \
  do {                                                                                     \
    const bool sw = (ck[y] > ck[x]) | ((ck[y] == ck[x]) & (cf[y] > cf[x]));                \
    const int k_ = sw ? ck[y] : ck[x], f_ = sw ? cf[y] : cf[x];                            \
    ck[y] = sw ? ck[x] : ck[y]; cf[y] = sw ? cf[x] : cf[y];                                \
    ck[x] = k_; cf[x] = f_;                                                                \
  } while (0)
    if constexpr (NP * K == 4) { K2B_CE2(0, 1); K2B_CE2(2, 3); K2B_CE2(0, 2); K2B_CE2(1, 3); K2B_CE2(1, 2); }
    if constexpr (NP * K == 2) { K2B_CE2(0, 1); }
#undef K2B_CE2
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int wk = __reduce_max_sync(full, ck[0]);
      const int wf = __reduce_max_sync(full, (ck[0] == wk) ? cf[0] : -1);
      const bool pop = (cf[0] == wf) & (wf >= 0);
#pragma unroll
      for (int i = 0; i + 1 < NP * K; ++i) { ck[i] = pop ? ck[i + 1] : ck[i]; cf[i] = pop ? cf[i + 1] : cf[i]; }
      ck[NP * K - 1] = pop ? kKeyNone : ck[NP * K - 1]; cf[NP * K - 1] = pop ? -1 : cf[NP * K - 1];
      my_v = (lane == r) ? funkey(wk) : my_v;
      my_f = (lane == r) ? wf : my_f;
    }
  } else {
#pragma unroll
    for (int r = 0; r < K; ++r) {
      int bk = ck[0], bf = cf[0];
#pragma unroll
      for (int i = 1; i < NP * K; ++i) {                      // predicated selects, no divergent branches
        const bool b = (ck[i] > bk) | ((ck[i] == bk) & (cf[i] > bf));
        bk = b ? ck[i] : bk; bf = b ? cf[i] : bf;
      }
      const int wk = __reduce_max_sync(full, bk);
      const int wf = __reduce_max_sync(full, (bk == wk) ? bf : -1);
#pragma unroll
      for (int i = 0; i < NP * K; ++i) {
        const bool hit = (cf[i] == wf) & (wf >= 0);
        cf[i] = hit ? -1 : cf[i]; ck[i] = hit ? kKeyNone : ck[i];
      }
      my_v = (lane == r) ? funkey(wk) : my_v;
      my_f = (lane == r) ? wf : my_f;
    }
  }

  K2B_SUB(10);
  // ---- lanes 0..K-1 hold the winners, best first -------------------------------------------------------------
  const bool cand = lane < K && my_f >= 0;
  int par = 0, tok = -1, c0 = -1, c1 = blank, ln = 2, cs = 0;
  uint64_t hs = kHashSeedC;
  if (cand) {
#pragma unroll
    for (int hh = 1; hh < K; ++hh) par += (my_f >= hh * V) ? 1 : 0;
    const int y = my_f - par * V;
    const int prow = s * K + par;
    hs = in.hash[prow]; ln = in.len[prow]; c0 = in.ctx0[prow]; c1 = in.ctx1[prow];
    if (CG) cs = in.cst[prow];
    if (y != blank && y != unk && y != extra_mask) {
      tok = y; hs = hash_push_c(hs, y); ln += 1; c0 = c1; c1 = y;
      if (CG) {                       // hot words: the boost goes to the extended hypothesis, after the top-K selection
        my_v += __ldg(cg_delta + (size_t)cs * V + y);
        cs = __ldg(cg_next + (size_t)cs * V + y);
      }
      if (do_prefetch)    // pull the new context's decoder row towards L2 while the merge finishes (one TMA-unit op)
        l2_prefetch_bulk(dec_tab + ((size_t)(c0 + 1) * V + c1) * J, (uint32_t)(J * 4));
    }
  }
  __syncwarp();
  K2B_SUB(11);
  int root = lane;
  float lp = my_v;
  if constexpr (K > 1) {
    // scratch row q: {hash lo, hash hi, length (negative = no candidate), ctx0 | ctx1, score, root, -, -}
    uint4* scr4 = reinterpret_cast<uint4*>(scr);
    if (lane < K) {
      scr4[2 * lane] = make_uint4((uint32_t)hs, (uint32_t)(hs >> 32), (uint32_t)(cand ? ln : -1 - lane), (uint32_t)c0);
      scr4[2 * lane + 1] = make_uint4((uint32_t)c1, __float_as_uint(my_v), 0u, 0u);
    }
    __syncwarp();
#pragma unroll
    for (int q = K - 1; q >= 0; --q) {            // descending: the lowest equal q wins
      const uint4 u = scr4[2 * q];
      const uint32_t uc1 = scr[8 * q + 4];
      const bool eq = (q < lane) & cand & (u.x == (uint32_t)hs) & (u.y == (uint32_t)(hs >> 32)) & (u.z == (uint32_t)ln) &
                      (u.w == (uint32_t)c0) & (uc1 == (uint32_t)c1);
      root = eq ? q : root;
    }
    K2B_SUB(12);
    const unsigned merged = __ballot_sync(full, cand && root != lane);
    if (merged) {                      // log-add merged scores into their root, in insertion (rank) order
      if (lane < K) scr[8 * lane + 6] = (uint32_t)root;
      __syncwarp();
#pragma unroll
      for (int q = 1; q < K; ++q) {
        const int qroot = (int)scr[8 * q + 6];
        const float qv = __uint_as_float(scr[8 * q + 5]);
        if (cand && ((merged >> q) & 1u) && qroot == lane) lp = logaddexp_c(lp, qv);
      }
    }
  }
  K2B_SUB(13);
  const bool is_root = cand && root == lane;
  const unsigned roots = __ballot_sync(full, is_root);
  const int nnew = __popc(roots);
  if (is_root) {
    const int slot = __popc(roots & ((1u << lane) - 1u));
    const int o = s * K + slot;
    out.ctx0[o] = c0; out.ctx1[o] = c1; out.lp[o] = lp; out.len[o] = ln; out.hash[o] = hs;
    if (CG) out.cst[o] = cs;
    if (bp_row != nullptr) bp_row[slot] = (par << 28) | (tok + 1);
  }
  if (lane >= nnew && lane < K) {
    const int o = s * K + lane;
    out.ctx0[o] = -1; out.ctx1[o] = blank; out.lp[o] = -INFINITY; out.len[o] = 2; out.hash[o] = kHashSeedC;
    if (CG) out.cst[o] = 0;
    if (bp_row != nullptr) bp_row[lane] = 0;
  }
  if (lane == 0) out.nlive[s] = nnew;
  __syncwarp();                       // the scratch is reused by this warp's next stream
  K2B_SUB(14);
#undef K2B_SUB
}

// Warp roles: warps 0..15 build the joiner operand (two hypothesis rows each), read the accumulator out, reduce and merge;
// warp 16 only issues MMAs (as the K-quarters of the operand land) - so the tensor pipe starts on the first quarter while
// the other three are still being computed.
//
// PAIR (even cluster sizes): the CTAs 2i, 2i+1 of the cluster form a tcgen05 CTA pair (cta_group::2, M = 256 = their two
// vocabulary slices). The B operand of a pair MMA is split by columns between the two CTAs, so each CTA builds only 16 of
// the 32 hypothesis rows (one per worker warp): half the decoder-row traffic, half the tanh / split work, half the stores.
// The even CTA's MMA warp issues for both (its bar_q mbarriers also count the peer's warps, arriving through DSMEM), the
// commit is multicast to both CTAs' bar_mma; everything after the accumulator read-out is unchanged.
//
// NTC >= 0 (J = 512 only, not PAIR): the first NTC of the 8 k-blocks of W_hi are held in TENSOR MEMORY (beside W_lo and the
// accumulator) and their products are TS MMAs; the MMA issue loop is fully unrolled with every descriptor a compile-time offset
// from two uniform registers (a run-time split costs ~65 cycles of dependent uniform arithmetic per k-block on the issue chain,
// more than it saves). NTC < 0: any J, everything of W_hi in shared memory.
template <int K, bool X3, bool TIMED, bool PAIR, int NTC>
__global__ void __maxnreg__(K2B_CLUSTER_MAXREG) cluster_beam_kernel(const ClusterArgs a) {
  static_assert(NTC < 0 || (!PAIR && NTC <= 8 && 64 + (X3 ? 256 : 0) + 32 * NTC <= 512), "W_hi k-blocks in tensor memory: no room");
  constexpr int XWP = xw_padded(K);
  constexpr int S = kNH / K;
  constexpr int kRowsCta = PAIR ? 16 : 32;            // hypothesis rows this CTA builds
  constexpr int kXTile = 2 * kRowsCta * 128;          // one k-block of the stacked [x_hi rows; x_lo rows] operand
  constexpr int RPW = PAIR ? 1 : 2;                   // rows built per worker warp
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ HypState st[2];
  __shared__ float bias_s[128];
  __shared__ uint64_t bar_w, bar_mma, bar_q[4];   // bar_q[i]: K-quarter i of the joiner operand is in shared memory
  __shared__ uint64_t xbar[2];                    // partials of frame t (buffer t & 1): 16 local warps + remote st.async bytes
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) uint32_t sel_scr[kWorkers][8 * K];   // per-warp winner exchange of the merge
  // back-pointer rows of the last kBpSteps frames (CTA 0 of the cluster): a global store per winner inside the merge cost ~100
  // cycles of every step (address arithmetic + the store on the merge warp's chain); every kBpSteps frames all worker warps write
  // the block out in 512-byte runs instead
  __shared__ int32_t bp_ring[kBpSteps][kNH];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);     // the compiler knows this one is warp-uniform
  const bool worker = warp_u < kWorkers;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x / a.CS;
  const int J = NTC >= 0 ? 512 : a.J, V = a.V, CS = a.CS, T = a.T;
  const int nkb = J / 64;
  const int nt = NTC >= 0 ? NTC : 0;                                  // W_hi k-blocks [0, nt) are read from tensor memory
  uint8_t* w_hi = smem;                                               // k-blocks [nt, nkb): tiles of 128 rows x 128 B
  uint8_t* xop = w_hi + (size_t)(nkb - nt) * 16384;                   // nkb tiles of 64 rows x 128 B
  float* Lt = reinterpret_cast<float*>(xop);                          // aliases the operand between MMA and next build
  float* xch = reinterpret_cast<float*>(xop + xop_region_bytes(nkb, kXTile, kLtStride));  // [2][CS][NH][XWP]

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_q[i], PAIR ? 2 * kWorkers : kWorkers);
    for (int i = 0; i < 2; ++i) mbar_init(&xbar[i], kWorkers);
    mbar_fence_init();
  }
  if (PAIR) cluster_sync();                       // both CTAs of a pair are resident before the pair-wide TMEM allocation
  if (warp == 0) { if (PAIR) tmem_alloc2(&tmem_slot, 512); else tmem_alloc(&tmem_slot, 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  // TMEM columns: accumulator 64 | (PAIR: second accumulator 32 for the W_lo product, whose column order differs) | W_lo J/2
  // (X3 only) | the first nt k-blocks of W_hi, 32 columns each
  const uint32_t t_d = tbase, t_d2 = tbase + 64, t_wlo = PAIR ? tbase + 128 : tbase + 64;
  const uint32_t t_whi = X3 ? t_wlo + (uint32_t)(J / 2) : t_wlo;
  const uint32_t prank = PAIR ? (rank & 1u) : 0u;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;

  // ---- one-time staging of this CTA's weight slice ------------------------------------------------------
  if (tid == 0) {
    mbar_expect_tx(&bar_w, (uint32_t)((nkb - nt) * 16384));
    for (int kb = nt; kb < nkb; ++kb)
      tma_bulk_g2s(w_hi + (size_t)(kb - nt) * 16384, a.wo_hi_img + ((size_t)rank * nkb + kb) * 16384, 16384, &bar_w);
  }
  // (every worker warp takes part: a warp reaches the 32 TMEM lanes of its quarter, the four warps of a quarter share the columns)
  if (nt > 0 && worker) {
    const uint32_t* src = a.wo_hi_rows + ((size_t)rank * 128 + 32 * (warp & 3) + lane) * (J / 2);
    for (int c0 = 32 * (warp >> 2); c0 < nt * 32; c0 += 32 * (kWorkers / 4)) {
      uint32_t v[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + c0) + q);
        v[4 * q] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
      }
      tmem_st32(t_whi + lane_base + (uint32_t)c0, v);
    }
    tmem_st_wait();
  }
  if (X3 && worker) {
    const uint32_t* src = a.wo_lo + ((size_t)rank * 128 + 32 * (warp & 3) + lane) * (J / 2);
    for (int c0 = 32 * (warp >> 2); c0 < J / 2; c0 += 32 * (kWorkers / 4)) {
      uint32_t v[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + c0) + q);
        v[4 * q] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
      }
      tmem_st32(t_wlo + lane_base + (uint32_t)c0, v);
    }
    tmem_st_wait();
  }
  if (tid < 128) bias_s[tid] = a.bias[rank * 128 + tid];
  if (tid < kNH) {
    const int n = tid, h = n % K;
    for (int b = 0; b < 2; ++b) {
      int c0i = -1, c1i = a.blank;
      const int gs = cluster * S + n / K;
      if (a.hyp_in != nullptr && h == 0 && gs < a.B) { c0i = (int)a.hyp_in[2 * gs]; c1i = (int)a.hyp_in[2 * gs + 1]; }
      st[b].ctx0[n] = c0i; st[b].ctx1[n] = c1i;
      st[b].lp[n] = (h == 0) ? 0.f : -INFINITY;
      st[b].len[n] = 2; st[b].hash[n] = kHashSeedC;
      st[b].nlive[n] = 0;
      st[b].cst[n] = 0;
    }
    if (n < S) {
      const int gs = cluster * S + n;
      st[0].nlive[n] = gs < a.B ? (a.resume ? a.fin_nlive[gs] : 1) : 0;
    }
    if (a.resume) {                       // continue where the previous time chunk stopped
      const int gs = cluster * S + n / K;
      if (gs < a.B) {
        const size_t gi = (size_t)gs * K + h;
        st[0].ctx0[n] = a.io_ctx[2 * gi]; st[0].ctx1[n] = a.io_ctx[2 * gi + 1];
        st[0].lp[n] = a.fin_lp[gi]; st[0].len[n] = a.fin_len[gi]; st[0].hash[n] = a.io_hash[gi];
        if (a.io_cst != nullptr) st[0].cst[n] = a.io_cst[gi];
      }
    }
  }
  bool ok = true;
  if (tid == 0) ok = mbar_wait(&bar_w, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync();            // every CTA of the cluster is resident: remote shared memory may be written

  __shared__ long long tph[20];          // phase cycle totals live in shared memory: no registers held across the loop
  const bool timed = TIMED && a.timing != nullptr && blockIdx.x == 0 && tid == 0;
  if (timed) for (int i = 0; i < 20; ++i) tph[i] = 0;
  long long tlast = timed ? clock64() : 0;
#define K2B_PHASE(i) do { if (TIMED && timed) { const long long now = clock64(); tph[i] += now - tlast; tlast = now; } } while (0)

  const int nq = J / 4;

  if (!worker) {
    // =============================== MMA warp =================================================================
    // D[128 vocab, hyps] = W_slice * x^T. X3: one SS MMA with the stacked operand (N = 64: cols 0-31 = Wh*xh, 32-63 = Wh*xl)
    // + one TS MMA (A = Wl resident in TMEM, N = 32) accumulating Wl*xh into cols 0-31. Warp-uniform loop: the descriptors
    // stay in uniform registers, one elected lane issues.
    const uint32_t idesc64 = umma_idesc_bf16_f32(PAIR ? 256 : 128, 64), idesc32 = umma_idesc_bf16_f32(PAIR ? 256 : 128, 32);
    const uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);
    const uint32_t w_lo0 = ((smem_u32(w_hi) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t x_lo0 = ((smem_u32(xop) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t el = elect_one();
    if (!PAIR || prank == 0) {
      // local index of the first frame of the next time chunk (a.t0 + that is a multiple of ready_len), or never
      int next_ready = a.ready != nullptr ? a.ready_len - a.t0 % a.ready_len : 0x7fffffff;
      for (int t = 0; t < T; ++t) {
        uint32_t acc = 0;
        if constexpr (NTC >= 0) {
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            if (!mbar_wait(&bar_q[qd], (uint32_t)(t & 1))) ok = false;
            tc_fence_after();
#pragma unroll
            for (int kb = 2 * qd; kb < 2 * qd + 2; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t dx = ((uint64_t)desc_hi << 32) | (uint64_t)(x_lo0 + (uint32_t)(kb * (kXTile >> 4) + k * 2));
                if (kb < NTC) {
                  umma_ts_e(t_d, t_whi + (uint32_t)((kb * 4 + k) * 8), dx, X3 ? idesc64 : idesc32, acc, el);
                } else {
                  const uint64_t dw = ((uint64_t)desc_hi << 32) | (uint64_t)(w_lo0 + (uint32_t)((kb - NTC) * (16384 >> 4) + k * 2));
                  umma_ss_e(t_d, dw, dx, X3 ? idesc64 : idesc32, acc, el);
                }
                if (X3) umma_ts_e(t_d, t_wlo + (uint32_t)((kb * 4 + k) * 8), dx, idesc32, 1, el);
                acc = 1;
              }
            }
          }
        } else {
          for (int qd = 0; qd < 4; ++qd) {
            if (PAIR) { if (!mbar_wait_cluster(&bar_q[qd], (uint32_t)(t & 1))) ok = false; }
            else { if (!mbar_wait(&bar_q[qd], (uint32_t)(t & 1))) ok = false; }
            tc_fence_after();
            for (int kb = 2 * qd; kb < 2 * qd + 2 && kb < nkb; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t dw = ((uint64_t)desc_hi << 32) | (uint64_t)(w_lo0 + (uint32_t)(kb * (16384 >> 4) + k * 2));
                const uint64_t dx = ((uint64_t)desc_hi << 32) | (uint64_t)(x_lo0 + (uint32_t)(kb * (kXTile >> 4) + k * 2));
                if (PAIR) {
                  // pair MMA: columns [0,16) Wh*xh(rows of CTA 0), [16,32) Wh*xl(CTA 0), [32,48) Wh*xh(CTA 1), [48,64) Wh*xl(CTA 1);
                  // the W_lo product (each CTA supplies its 16 hi rows) goes to its own 32 columns
                  umma2_ss_e(t_d, dw, dx, X3 ? idesc64 : idesc32, acc, el);
                  if (X3) umma2_ts_e(t_d2, t_wlo + (uint32_t)((kb * 4 + k) * 8), dx, idesc32, acc, el);
                } else {
                  umma_ss_e(t_d, dw, dx, X3 ? idesc64 : idesc32, acc, el);
                  if (X3) umma_ts_e(t_d, t_wlo + (uint32_t)((kb * 4 + k) * 8), dx, idesc32, 1, el);
                }
                acc = 1;
              }
            }
          }
        }
        if (t + 2 == next_ready) {          // frame t + 2 is the first of a time chunk that may still be on its way
          if (next_ready < T && !wait_frames_ready(a.ready + (a.t0 + next_ready) / a.ready_len, a.ready_epoch)) ok = false;
          next_ready += a.ready_len;
        }
        if (PAIR) umma2_commit_e(&bar_mma, (uint16_t)(3u << (rank & ~1u)), el);
        else umma_commit_e(&bar_mma, el);
      }
    }
  } else {
    // =============================== worker warps ===============================================================
    // the two hypothesis rows of this warp belong to one stream when K is even (two streams for K = 1); the frame rows
    // are prefetched one step ahead
    constexpr int NE = (K % 2 == 0 || PAIR) ? 1 : 2;
    const int n0 = warp * 2;                                     // hypotheses this warp reduces / exchanges
    const int nb0 = PAIR ? (int)(16 * prank) + warp : warp * 2;  // first hypothesis row this warp builds (RPW of them)
    const int lr0 = PAIR ? warp : warp * 2;                      // ... and its row in this CTA's operand tile
    int g_w = cluster * S + nb0 / K;
    if (g_w >= a.B) g_w = a.B - 1;
    const float4* enc_row = reinterpret_cast<const float4*>(a.encE + ((size_t)g_w * a.Ttot + a.t0) * J);
    int g_w1 = cluster * S + (nb0 + RPW - 1) / K;
    if (g_w1 >= a.B) g_w1 = a.B - 1;
    const float4* enc_row1 = reinterpret_cast<const float4*>(a.encE + ((size_t)g_w1 * a.Ttot + a.t0) * J);
    float4 ecur[4], ecur1[NE == 2 ? 4 : 1];
    uint32_t xoff[RPW][4];         // loop-invariant swizzled byte offsets of this thread's operand chunks (hi rows)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = lane + 32 * i;
      if (q < nq) ecur[i] = __ldg(enc_row + q);
      if (NE == 2 && q < nq) ecur1[NE == 2 ? i : 0] = __ldg(enc_row1 + q);
      const int k = 4 * q;
#pragma unroll
      for (int r = 0; r < RPW; ++r) xoff[r][i] = (uint32_t)(k >> 6) * kXTile + sw128_offset(lr0 + r, k & 63);
    }
    const int col0 = 8 * (warp >> 2);            // this warp's 8 accumulator columns (hypotheses) in the read-out
    const int vrow = 32 * (warp & 3) + lane;     // ... and its vocabulary row of the slice
    const float bsv = bias_s[vrow];
    const int nvalid = min(128, V - (int)rank * 128);   // vocabulary entries of this slice
    int cur = 0;

    for (int t = 0; t < T; ++t) {
      // ---- (a) joiner prologue: x[n,:] = tanh(enc[stream(n),t,:] + dec(ctx(n))) as bf16 hi/lo, K-major swizzled,
      //      K-quarter by K-quarter (each warp arrives on bar_q[i] after its part of quarter i)
      {
        const HypState& sc = st[cur];
        const float4* pd[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
          pd[r] = reinterpret_cast<const float4*>(a.dec_tab + ((size_t)(sc.ctx0[nb0 + r] + 1) * V + sc.ctx1[nb0 + r]) * J);
        // fence.proxy.async waits for every outstanding load of the thread (MEMBAR.ALL.CTA), so only the first quarter's
        // decoder-row loads are issued before the first fence: the tensor pipe starts after one L2 round trip for 1/4 of
        // the rows, and the other three quarters' loads (issued right after that fence) land while it works
        float4 dd[RPW][4];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
          if (lane < nq) dd[r][0] = __ldg(pd[r] + lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int q = lane + 32 * i;
          if (i == 1) {
#pragma unroll
            for (int i2 = 1; i2 < 4; ++i2) {
              const int q2 = lane + 32 * i2;
#pragma unroll
              for (int r = 0; r < RPW; ++r)
                if (q2 < nq) dd[r][i2] = __ldg(pd[r] + q2);
            }
          }
          if (q < nq) {
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
              const float4 d = dd[r][i];
              const float4 ev = (NE == 2 && r) ? ecur1[NE == 2 ? i : 0] : ecur[i];
              const float x0 = tanh_from_exp(ev.x, d.x), x1 = tanh_from_exp(ev.y, d.y);
              const float x2 = tanh_from_exp(ev.z, d.z), x3 = tanh_from_exp(ev.w, d.w);
              uint8_t* dst = xop + xoff[r][i];
              if (X3) {
                // split by truncation: hi = upper 16 bits (one PRMT per pair), lo = x - hi exactly, rounded to bf16
                const uint32_t b0 = __float_as_uint(x0), b1 = __float_as_uint(x1), b2 = __float_as_uint(x2), b3 = __float_as_uint(x3);
                *reinterpret_cast<uint2*>(dst) = make_uint2(__byte_perm(b0, b1, 0x7632), __byte_perm(b2, b3, 0x7632));
                const float l0 = x0 - __uint_as_float(b0 & 0xffff0000u), l1 = x1 - __uint_as_float(b1 & 0xffff0000u);
                const float l2 = x2 - __uint_as_float(b2 & 0xffff0000u), l3 = x3 - __uint_as_float(b3 & 0xffff0000u);
                *reinterpret_cast<uint2*>(dst + kRowsCta * 128) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
              } else {
                *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(x0, x1), pack_bf16x2(x2, x3));
              }
            }
          }
          fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async proxy
          __syncwarp();
          if (lane == 0) {
            if (PAIR && prank) mbar_arrive_remote(dsmem_map(smem_u32(&bar_q[i]), rank ^ 1u));   // the pair's leader issues the MMAs
            else mbar_arrive(&bar_q[i]);
          }
          __syncwarp();                      // reconverge: without it the compiler may leave lane 0 split off for the rest of the step
          K2B_PHASE(15 + i);
        }
        if (t + 1 < T) {
          const float4* pe = enc_row + (size_t)(t + 1) * nq;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int q = lane + 32 * i;
            if (q < nq) ecur[i] = __ldg(pe + q);
            if (NE == 2 && q < nq) ecur1[NE == 2 ? i : 0] = __ldg(enc_row1 + (size_t)(t + 1) * nq + q);
          }
        }
      }
      K2B_PHASE(0);

      // ---- (c) accumulator -> registers (+bias) -> transposed shared tile; every warp reads 8 columns of its lane quarter
      if (lane == 0 && !mbar_wait(&bar_mma, (uint32_t)(t & 1))) ok = false;
      __syncwarp();
      tc_fence_after();
      K2B_PHASE(1);
      {
        uint32_t va[8], vb[8], vc[8];
        if (PAIR) {
          // hypothesis h = 16 p + j was built by CTA p of the pair: Wh*xh in column 32 p + j, Wh*xl in 32 p + 16 + j, Wl*xh in
          // column 16 p + j of the second accumulator (without X3: Wh*xh in column 16 p + j)
          const uint32_t pp = (uint32_t)(col0 >> 4), j0 = (uint32_t)(col0 & 15);
          tmem_ld8(t_d + lane_base + (X3 ? 32u * pp + j0 : 16u * pp + j0), va);
          if (X3) {
            tmem_ld8(t_d + lane_base + 32u * pp + 16u + j0, vb);
            tmem_ld8(t_d2 + lane_base + 16u * pp + j0, vc);
          }
        } else {
          tmem_ld8(t_d + lane_base + (uint32_t)col0, va);
          if (X3) tmem_ld8(t_d + lane_base + (uint32_t)(32 + col0), vb);
        }
        tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          float v = __uint_as_float(va[n]) + bsv;
          if (X3) v += __uint_as_float(vb[n]);
          if (X3 && PAIR) v += __uint_as_float(vc[n]);
          Lt[vrow * kLtStride + col0 + n] = v;
        }
        tc_fence_before();
      }
      named_bar_sync(1, kWorkers * 32);
      K2B_PHASE(2);

      // ---- (d) per hypothesis (two per warp, interleaved): sum-exp and top-K over this slice's 128 logits. Selection keys
      //      are the order-preserving integer image of the logit with its low 7 bits replaced by the position in the slice:
      //      unique, so one REDUX per round finds value and position at once (ties and values closer than 2^-16 relative
      //      resolve to the larger vocabulary index; the exact fp32 logit is re-read for the record).
      float* xw = xch + (size_t)(t & 1) * CS * kNH * XWP;
      {
        float v[2][4];
        int pk[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int pos = lane + 32 * j;
            v[r][j] = Lt[pos * kLtStride + n0 + r];
            const int kb = __float_as_int(v[r][j]);
            const int key = kb ^ ((kb >> 31) & 0x7fffffff);
            pk[r][j] = (key & ~127) | pos;      // rows beyond V carry a -inf bias: their keys lose against every valid one
          }
          // sort the four unique keys of this lane once, best first: every round then pops the head
#define K2B_CE(x, y) do { const int hi_ = max(pk[r][x], pk[r][y]), lo_ = min(pk[r][x], pk[r][y]); pk[r][x] = hi_; pk[r][y] = lo_; } while (0)
          K2B_CE(0, 1); K2B_CE(2, 3); K2B_CE(0, 2); K2B_CE(1, 3); K2B_CE(1, 2);
#undef K2B_CE
        }
        int keep[2] = {kKeyNone, kKeyNone};
        float m[2], sum[2];
#pragma unroll
        for (int rr = 0; rr < K; ++rr) {
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int wk = __reduce_max_sync(0xffffffffu, pk[r][0]);
            const bool won = pk[r][0] == wk;
            pk[r][0] = won ? pk[r][1] : pk[r][0];
            pk[r][1] = won ? pk[r][2] : pk[r][1];
            pk[r][2] = won ? pk[r][3] : pk[r][2];
            pk[r][3] = won ? kKeyNone : pk[r][3];
            keep[r] = (lane == rr) ? wk : keep[r];
            if (rr == 0) {
              // softmax offset = the (truncated, hence <=) slice maximum; fixed-point sum: order-independent, one REDUX
              const int mk = wk & ~127;
              m[r] = (wk == kKeyNone) ? -INFINITY : __int_as_float(mk ^ ((mk >> 31) & 0x7fffffff));
              sum[r] = 1.f;
              if (K > 1 || a.need_lp) {
                const float mneg = (wk == kKeyNone) ? 0.f : -m[r] * 1.4426950408889634f;
                float ls = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float ex = ex2_approx(fmaf(v[r][j], 1.4426950408889634f, mneg));
                  ls += ex;                      // (exp2(-inf) = 0 for the rows beyond V)
                }
                const unsigned tot = __reduce_add_sync(0xffffffffu, __float2uint_rn(ls * 16777216.f));
                sum[r] = (wk == kKeyNone) ? 0.f : (float)tot * (1.f / 16777216.f);
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float* mine = xw + ((size_t)rank * kNH + n0 + r) * XWP;
          if (lane == 0) *reinterpret_cast<float2*>(mine) = make_float2(m[r], sum[r]);
          if (lane < K) {
            const int pos = keep[r] & 127;
            const bool has = keep[r] != kKeyNone && pos < nvalid;      // a slice with fewer than K valid rows
            mine[2 + lane] = has ? Lt[pos * kLtStride + n0 + r] : -INFINITY;
            reinterpret_cast<int*>(mine)[2 + K + lane] = has ? (int)rank * 128 + pos : -1;
          }
        }
        __syncwarp();
        K2B_PHASE(3);
        // both records (contiguous) go into the same slots of every other CTA of the cluster: 16-byte asynchronous
        // stores that complete on the destination's mbarrier
        constexpr int kChunks = 2 * XWP / 4;
        const uint32_t base = smem_u32(xw + ((size_t)rank * kNH + n0) * XWP);
        const uint32_t xb_addr = smem_u32(&xbar[t & 1]);
        for (int it = lane; it < kChunks * (CS - 1); it += 32) {
          const int dsel = it / kChunks, ch = it - dsel * kChunks;
          uint32_t dst = rank + 1 + (uint32_t)dsel;
          if (dst >= (uint32_t)CS) dst -= (uint32_t)CS;
          const uint4 q = lds_v4(base + 16u * ch);
          dsmem_st_async_v4(dsmem_map(base + 16u * ch, dst), q.x, q.y, q.z, q.w, dsmem_map(xb_addr, dst));
        }
        if (lane == 0) {
          if (warp == 0) mbar_expect_tx(&xbar[t & 1], (uint32_t)((CS - 1) * kNH * XWP * 4));
          else mbar_arrive(&xbar[t & 1]);
        }
        __syncwarp();                        // reconverge (see above)
      }
      K2B_PHASE(4);

      // ---- (e) hypothesis merge, redundantly in every CTA (deterministic, so no second exchange) ---------------------
      if (warp < S) {
        if (!mbar_wait(&xbar[t & 1], (uint32_t)((t >> 1) & 1))) ok = false;
        K2B_PHASE(5);
        for (int s = warp; s < S; s += kWorkers) {
          const int g = cluster * S + s;
          int32_t* bp_row = (rank == 0) ? &bp_ring[t & (kBpSteps - 1)][s * K] : nullptr;
          if (a.lens != nullptr && g < a.B && a.t0 + t >= a.lens[g]) {
            // past the end of this stream (ragged batch): the hypotheses stay as they are, the back-pointers are the identity
            if (lane < K) {
              const int o = s * K + lane;
              st[cur ^ 1].ctx0[o] = st[cur].ctx0[o]; st[cur ^ 1].ctx1[o] = st[cur].ctx1[o]; st[cur ^ 1].lp[o] = st[cur].lp[o];
              st[cur ^ 1].len[o] = st[cur].len[o]; st[cur ^ 1].hash[o] = st[cur].hash[o];
              if (a.cg_next != nullptr) st[cur ^ 1].cst[o] = st[cur].cst[o];
              if (bp_row != nullptr) bp_row[lane] = lane < st[cur].nlive[s] ? (lane << 28) : 0;
            }
            if (lane == 0) st[cur ^ 1].nlive[s] = st[cur].nlive[s];
            __syncwarp();
            continue;
          }
          if (K > 1 && a.cg_next != nullptr)
            select_stream<K, true>(s, V, CS, xw, st[cur], st[cur ^ 1], a.blank, a.unk, a.extra_mask, bp_row, lane, a.dec_tab, J,
                                   (int)rank == (s % CS), sel_scr[warp], a.need_lp != 0, a.cg_next, a.cg_delta, nullptr);
          else
            select_stream<K, false>(s, V, CS, xw, st[cur], st[cur ^ 1], a.blank, a.unk, a.extra_mask, bp_row, lane, a.dec_tab, J,
                                    (int)rank == (s % CS), sel_scr[warp], a.need_lp != 0, nullptr, nullptr, (TIMED && timed) ? tph : nullptr);
        }
      }
      K2B_PHASE(6);
      named_bar_sync(1, kWorkers * 32);
      cur ^= 1;
      if (rank == 0 && ((t & (kBpSteps - 1)) == kBpSteps - 1 || t == T - 1)) {
        // bp [B, Ttot, K]: the (step, slot) entries of one stream are contiguous
        const int t_first = t & ~(kBpSteps - 1), per = (t - t_first + 1) * K;
        for (int i = tid; i < S * per; i += kWorkers * 32) {
          const int s = i / per, rem = i - s * per, step = rem / K, k = rem - step * K;
          const int g = cluster * S + s;
          if (g < a.B) a.bp[((size_t)g * a.Ttot + a.t0 + t_first + step) * K + k] = bp_ring[step][s * K + k];
        }
      }
      K2B_PHASE(7);
    }
    if (timed) for (int i = 0; i < 20; ++i) a.timing[i] = tph[i];

    if (rank == 0 && tid < S * K) {
      const int s = tid / K, hslot = tid % K, g = cluster * S + s;
      if (g < a.B) {
        a.fin_lp[(size_t)g * K + hslot] = st[cur].lp[tid];
        a.fin_len[(size_t)g * K + hslot] = st[cur].len[tid];
        if (a.io_cst != nullptr) a.io_cst[(size_t)g * K + hslot] = st[cur].cst[tid];
        if (a.io_ctx != nullptr) {
          a.io_ctx[2 * ((size_t)g * K + hslot)] = st[cur].ctx0[tid];
          a.io_ctx[2 * ((size_t)g * K + hslot) + 1] = st[cur].ctx1[tid];
          a.io_hash[(size_t)g * K + hslot] = st[cur].hash[tid];
        }
        if (hslot == 0) a.fin_nlive[g] = st[cur].nlive[s];
        if (hslot == 0 && a.hyp_out != nullptr) {          // ref OnlineRecognizer.cs:208: last ctx tokens back into stream.Hyp
          a.hyp_out[2 * g] = st[cur].ctx0[tid];
          a.hyp_out[2 * g + 1] = st[cur].ctx1[tid];
        }
      }
    }
  }
#undef K2B_PHASE
  if (!ok) atomicExch(a.status, 1);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) { if (PAIR) tmem_dealloc2(tbase, 512); else tmem_dealloc(tbase, 512); }
}

// ---- weight-load-time packing ----------------------------------------------------------------------------------
__global__ void pack_out_w_kernel(const float* __restrict__ out_w, const float* __restrict__ out_b, int V, int J, int CS,
                                  uint8_t* __restrict__ img, uint32_t* __restrict__ lo, uint32_t* __restrict__ hi_rows,
                                  float* __restrict__ bias) {
  const int row = blockIdx.x;            // 0 .. CS*128-1
  const int c = row >> 7, r = row & 127, nkb = J / 64;
  if (threadIdx.x == 0) bias[row] = row < V ? out_b[row] : -INFINITY;
  for (int k = 2 * threadIdx.x; k < J; k += 2 * blockDim.x) {
    float x0 = 0.f, x1 = 0.f;
    if (row < V) { x0 = out_w[(size_t)row * J + k]; x1 = out_w[(size_t)row * J + k + 1]; }
    const float h0 = ptx::bf16_round(x0), h1 = ptx::bf16_round(x1);
    const size_t off = ((size_t)c * nkb + (k >> 6)) * 16384 + ptx::sw128_offset(r, k & 63);
    *reinterpret_cast<uint32_t*>(img + off) = ptx::pack_bf16x2(h0, h1);
    lo[(size_t)row * (J / 2) + (k >> 1)] = ptx::pack_bf16x2(x0 - h0, x1 - h1);
    hi_rows[(size_t)row * (J / 2) + (k >> 1)] = ptx::pack_bf16x2(h0, h1);
  }
}

__global__ void enum_ctx_kernel(int V, long long first, int n, int32_t* __restrict__ ctx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long id = first + i;
  ctx[2 * i] = (int)(id / V) - 1;
  ctx[2 * i + 1] = (int)(id % V);
}

__global__ void exp2x_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = expf(2.f * fminf(fmaxf(in[i], -21.f), 21.f));
}
// time chunk [t0, t0 + tc) of every stream: in [B,tc,J] compact -> out [B,T,J]
__global__ void exp2x_chunk_kernel(const float4* __restrict__ in, float4* __restrict__ out, int B, int tc, int T, int t0, int J4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * tc * J4) return;
  const int j = (int)(i % J4);
  const size_t bt = i / J4;
  const int t = (int)(bt % tc), b = (int)(bt / tc);
  const float4 x = __ldg(in + i);
  float4 y;
  y.x = expf(2.f * fminf(fmaxf(x.x, -21.f), 21.f)); y.y = expf(2.f * fminf(fmaxf(x.y, -21.f), 21.f));
  y.z = expf(2.f * fminf(fmaxf(x.z, -21.f), 21.f)); y.w = expf(2.f * fminf(fmaxf(x.w, -21.f), 21.f));
  out[((size_t)b * T + t0 + t) * J4 + j] = y;
}

}  // namespace

// `timed` selects the instrumented build of the beam-4 kernel (per-phase clock64 totals); production launches never carry it.
// `pair`: CTA-pair variant (even cluster sizes), opt-in with K2B_PAIR=1: measured on cfg2 it halves the prologue (quarters
// 1053 + 742 + 368 + 325 cycles against 1169 + 1028 + 474 + 438) but a cta_group::2 MMA at N <= 64 costs ~70 cycles against
// ~57 for the single-CTA form, so the step gets longer (1449 us per launch against 1383 us).
template <bool X3, bool PAIR, int NTC>
static void (*cluster_kernel_kx(int K))(const ClusterArgs) {
  return K == 1 ? cluster_beam_kernel<1, X3, false, PAIR, NTC> : (K == 2 ? cluster_beam_kernel<2, X3, false, PAIR, NTC>
       : (K == 4 ? cluster_beam_kernel<4, X3, false, PAIR, NTC> : cluster_beam_kernel<8, X3, false, PAIR, NTC>));
}
// W_hi k-blocks in tensor memory (J = 512): as many as fit beside the accumulator and W_lo
constexpr int kNtX3 = 6, kNtBf16 = 8;
// `nt` > 0 selects the instantiation with that many W_hi k-blocks in tensor memory (cluster_plan only offers kNtX3 / kNtBf16)
static void (*cluster_kernel_for(int K, bool x3, bool timed, bool pair, int nt))(const ClusterArgs) {
  if (timed && K == 4) {
    if (pair) return x3 ? cluster_beam_kernel<4, true, true, true, -1> : cluster_beam_kernel<4, false, true, true, -1>;
    if (nt > 0) return x3 ? cluster_beam_kernel<4, true, true, false, kNtX3> : cluster_beam_kernel<4, false, true, false, kNtBf16>;
    return x3 ? cluster_beam_kernel<4, true, true, false, -1> : cluster_beam_kernel<4, false, true, false, -1>;
  }
  if (timed && K == 1 && x3 && !pair && nt == 0) return cluster_beam_kernel<1, true, true, false, -1>;
  if (pair) return x3 ? cluster_kernel_kx<true, true, -1>(K) : cluster_kernel_kx<false, true, -1>(K);
  if (nt > 0) return x3 ? cluster_kernel_kx<true, false, kNtX3>(K) : cluster_kernel_kx<false, false, kNtBf16>(K);
  return x3 ? cluster_kernel_kx<true, false, -1>(K) : cluster_kernel_kx<false, false, -1>(K);
}
static bool cluster_pair_mode(const k2b_handle* h, int CS) { return CS % 2 == 0 && h->opt_pair != 0; }

static size_t cluster_dyn_smem(int J, int CS, int K, int nt = 0) {
  const int nkb = J / 64;
  return (size_t)(nkb - nt) * 16384 + xop_region_bytes(nkb, 64 * 128, kLtStride) + 2ull * CS * kNH * xw_padded(K) * 4;
}

// Where the A operand lives. Measured on cfg2 (tools/time_cluster_variants.py, profiles/r02_cluster_placement.txt): every k-block
// of W_hi that is read from tensor memory instead of shared memory (a TS MMA instead of an SS one) takes ~130 cycles off the 64-MMA
// issue chain of a frame step, so at J = 512 as many as fit beside the accumulator and W_lo go there (6 of 8 with split-bf16 x3,
// all 8 in single-pass bf16) - and the shared memory they free lets V up to 1024 with beam 8 fit.
// k2b_set_option("wh_tmem_kb", 0) keeps everything in shared memory (comparison runs); other joiner widths always do.
struct ClusterPlan { int nt; size_t dyn; };
static ClusterPlan cluster_plan(const k2b_handle* h, int K, bool pair) {
  const k2b_config& c = h->cfg;
  const int J = c.joiner_dim, CS = (c.vocab_size + 127) / 128;
  const bool x3 = c.precision == K2B_PREC_BF16X3;
  int nt = (J == 512 && !pair && h->opt_wh_tmem != 0) ? (x3 ? kNtX3 : kNtBf16) : 0;
  return ClusterPlan{nt, cluster_dyn_smem(J, CS, K, nt)};
}

// V <= 1024: portable clusters of up to 8 CTAs, beams 1/2/4/8. 1024 < V <= 2048: 16-CTA (non-portable) clusters, greedy
// (beam 1) only, and only if the device can co-schedule such a cluster (occupancy query, cached on the handle).
bool cluster_path_supported(const k2b_handle* h, int K) {
  const k2b_config& c = h->cfg;
  const int CS = (c.vocab_size + 127) / 128;
  if (c.vocab_size > 2048 || c.joiner_dim > 512 || c.joiner_dim % 64) return false;
  if (K != 1 && K != 2 && K != 4 && K != 8) return false;
  if (CS > 8 && K != 1) return false;
  const size_t dyn = cluster_plan(h, K, cluster_pair_mode(h, CS)).dyn;
  if (dyn + 12288 > 227 * 1024) return false;      // + the kernel's static shared memory (7 - 11.5 KB)
  const size_t tab = (size_t)(c.vocab_size + 1) * c.vocab_size * c.joiner_dim * sizeof(float);
  if (tab > ((size_t)16 << 30)) return false;
  if (CS > 8) {
    k2b_handle* hm = const_cast<k2b_handle*>(h);
    if (hm->cluster16_ok < 0) {
      hm->cluster16_ok = 0;
      auto kern = cluster_kernel_for(1, c.precision == K2B_PREC_BF16X3, false, cluster_pair_mode(h, CS),
                                     cluster_plan(h, 1, cluster_pair_mode(h, CS)).nt);
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)CS); cfg.blockDim = dim3(kCAll); cfg.dynamicSmemBytes = dyn;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) == cudaSuccess && nc > 0) hm->cluster16_ok = 1;
      }
      cudaGetLastError();   // a failed query must not poison later launches
    }
    if (hm->cluster16_ok != 1) return false;
  }
  // the kernel reads the memoised decoder table: when that does not fit beside what else lives on the device (another handle, the
  // encoder's own session) the callers take the per-frame path, which needs none
  k2b_handle* hm = const_cast<k2b_handle*>(h);
  if (hm->weights_loaded && hm->dec_tab_state == 0) {
    bool have = false;
    if (ensure_dec_table(hm, &have) != K2B_OK) return false;
  }
  return hm->dec_tab_state > 0;
}

// The memoised stateless decoder: dec_tab[(y0+1)*V + y1] = exp(2*clamp(decoder(y0, y1))) for every context, J floats per row
// (513 MB at V = 500, 8.2 GB at V = 2000, 62.8 GB at V = 5537 - HBM capacity traded for a [N,D]x[D,J] GEMM per frame).
// Built once per weight load if it fits: at most 96 GiB and at most 60 % of the free device memory. *have = false otherwise
// (callers then run the decoder GEMM every frame).
int32_t ensure_dec_table(k2b_handle* h, bool* have) {
  *have = h->dec_tab != nullptr;
  if (h->dec_tab != nullptr || h->dec_tab_state < 0) return K2B_OK;
  const k2b_config& c = h->cfg;
  const int V = c.vocab_size, J = c.joiner_dim, D = c.decoder_dim;
  const long long nctx = (long long)(V + 1) * V;
  const size_t bytes = (size_t)nctx * J * sizeof(float);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
  if (bytes > ((size_t)96 << 30) || (double)bytes > 0.6 * (double)free_b ||
      cudaMalloc(reinterpret_cast<void**>(&h->dec_tab), bytes) != cudaSuccess) {
    cudaGetLastError();
    h->dec_tab = nullptr;
    h->dec_tab_state = -1;
    return K2B_OK;
  }
  const int chunk = 1 << 18;
  int32_t* ctx = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  auto build = [&]() -> int32_t {
    K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&ctx), sizeof(int32_t) * 2 * chunk));
    K2B_CUDA(h, cudaEventCreate(&ev0));
    K2B_CUDA(h, cudaEventCreate(&ev1));
    K2B_CUDA(h, cudaEventRecord(ev0, h->stream));
    for (long long first = 0; first < nctx; first += chunk) {
      const int n = (int)((nctx - first) < chunk ? (nctx - first) : chunk);
      enum_ctx_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(V, first, n, ctx);
      K2B_LAUNCH_CHECK(h);
      GemmArgs g;
      g.M = n; g.N = J; g.K = D;
      g.W = h->dec_w; g.bias = h->dec_b;
      g.ctx = ctx; g.tab0 = h->tab0; g.tab1 = h->tab1; g.V = V; g.neg_wrap = c.neg_id_mode == K2B_NEGID_WRAP; g.blank = c.blank_id;
      g.C = h->dec_tab + (size_t)first * J;
      K2B_TRY(launch_gemm_simt(h, PRO_DEC, EPI_EXP2X, g));
    }
    K2B_CUDA(h, cudaEventRecord(ev1, h->stream));
    K2B_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) h->dec_tab_build_ms = ms;
    return K2B_OK;
  };
  const int32_t st = build();
  if (ctx) cudaFree(ctx);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (st != K2B_OK) {                 // nothing half-built survives an error
    cudaFree(h->dec_tab);
    h->dec_tab = nullptr;
    h->dec_tab_state = -1;
    return st;
  }
  h->dec_tab_bytes = bytes;
  h->dec_tab_state = 1;
  *have = true;
  return K2B_OK;
}

// Builds (once per weight load) the shared-memory image / TMEM source / padded bias of the joiner weight and the
// memoised decoder table.
int32_t ensure_cluster_assets(k2b_handle* h) {
  if (h->tc_ready) return K2B_OK;
  const k2b_config& c = h->cfg;
  const int V = c.vocab_size, J = c.joiner_dim, D = c.decoder_dim, CS = (V + 127) / 128;
  const size_t rows = (size_t)CS * 128;
  bool have = false;
  K2B_TRY(ensure_dec_table(h, &have));      // first: without the table there is no cluster search and nothing else is allocated
  if (!have) return fail(h, K2B_ERR_STATE, "cluster search: the memoised decoder table could not be allocated");
  // (a retry after a failed allocation below re-uses what the earlier attempt obtained)
  if (h->wo_hi_img == nullptr) K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->wo_hi_img), rows * J * 2));
  if (h->wo_lo == nullptr) K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->wo_lo), rows * J * 2));
  if (h->wo_hi_rows == nullptr) K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->wo_hi_rows), rows * J * 2));
  if (h->bias_pad == nullptr) K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->bias_pad), rows * sizeof(float)));
  pack_out_w_kernel<<<(unsigned)rows, 128, 0, h->stream>>>(h->out_w, h->out_b, V, J, CS, h->wo_hi_img, h->wo_lo, h->wo_hi_rows, h->bias_pad);
  K2B_LAUNCH_CHECK(h);
  h->tc_ready = true;
  return K2B_OK;
}

int32_t exp2x_frames_chunk(k2b_handle* h, const float* in, float* out, int B, int tc, int T, int t0) {
  const int J4 = h->cfg.joiner_dim / 4;
  const size_t n = (size_t)B * tc * J4;
  exp2x_chunk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out),
                                                                         B, tc, T, t0, J4);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

int32_t exp2x_frames(k2b_handle* h, const float* in, float* out, size_t n) {
  exp2x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(in, out, n);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

// encE: [B,T,J] frames already mapped through exp(2x). Writes bp + final state; the caller runs the back-trace.
int32_t beam_cluster_dev(k2b_handle* h, const float* encE, int B, int T, int K, int32_t* bp, float* fin_lp, int32_t* fin_len,
                         int32_t* fin_nlive, int extra_mask, const int64_t* hyp_in, int64_t* hyp_out, int t0, int Ttot, int resume,
                         int32_t* io_ctx, unsigned long long* io_hash, bool need_lp, const int* ready, int ready_epoch,
                         int ready_len) {
  const k2b_config& c = h->cfg;
  const int V = c.vocab_size, J = c.joiner_dim, CS = (V + 127) / 128;
  const int S = kNH / K;
  const int nclusters = (B + S - 1) / S;
  int* status = h->dev_status;
  ClusterArgs a;
  a.ready = ready; a.ready_epoch = ready_epoch; a.ready_len = ready_len > 0 ? ready_len : 1;
  a.encE = encE; a.dec_tab = h->dec_tab; a.wo_hi_img = h->wo_hi_img; a.wo_lo = h->wo_lo; a.bias = h->bias_pad;
  a.B = B; a.T = T; a.K = K; a.V = V; a.J = J; a.S = S; a.CS = CS; a.blank = c.blank_id; a.unk = c.unk_id;
  a.x3 = c.precision == K2B_PREC_BF16X3 ? 1 : 0;
  a.extra_mask = extra_mask; a.hyp_in = hyp_in; a.hyp_out = hyp_out; a.need_lp = need_lp ? 1 : 0;
  a.t0 = t0; a.Ttot = Ttot > 0 ? Ttot : T; a.resume = resume; a.io_ctx = io_ctx; a.io_hash = io_hash;
  a.timing = h->cluster_timing;
  a.lens = h->lens_active ? h->lens_dev : nullptr;
  a.cg_next = nullptr; a.cg_delta = nullptr; a.io_cst = nullptr;
  if (h->cg_next != nullptr && K > 1) {               // greedy search (beam 1) is never biased
    a.cg_next = h->cg_next; a.cg_delta = h->cg_delta;
    a.io_cst = cluster_cst(h, B, K);
    if (a.io_cst == nullptr) return K2B_ERR_CUDA;
  }
  a.bp = bp; a.fin_lp = fin_lp; a.fin_len = fin_len; a.fin_nlive = fin_nlive; a.status = status;
  const ClusterPlan plan = cluster_plan(h, K, cluster_pair_mode(h, CS));
  a.wo_hi_rows = h->wo_hi_rows; a.nt = plan.nt;
  const size_t dyn = plan.dyn;
  void (*kern)(const ClusterArgs) = cluster_kernel_for(K, a.x3 != 0, a.timing != nullptr, cluster_pair_mode(h, CS), plan.nt);
  if (CS > 8) K2B_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  K2B_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nclusters * CS));
  cfg.blockDim = dim3(kCAll);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  prof_begin(h);
  K2B_CUDA(h, cudaLaunchKernelEx(&cfg, kern, a));
  prof_end(h);
  h->launches++;
  return K2B_OK;
}

int cluster_grid_ctas(const k2b_handle* h, int B, int K) {
  const int CS = (h->cfg.vocab_size + 127) / 128, S = kNH / K;
  return ((B + S - 1) / S) * CS;
}

namespace {
__global__ void set_ready_kernel(int* flag, int epoch) {
  if (threadIdx.x == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}
}  // namespace

int32_t cluster_set_ready(k2b_handle* h, int* flag, int epoch) {
  set_ready_kernel<<<1, 32, 0, h->stream>>>(flag, epoch);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

// dev_status: [0] cluster / persistent search kernels: an mbarrier or counter wait timed out, [1] the same in the tcgen05 GEMMs,
// [2] a caller-supplied Hyp held a token id outside the vocabulary (device-pointer calls), [3] scratch (first-emission frame)
// automaton state per hypothesis of the cluster engine: lives on the handle so that it is carried from one time chunk's launch
// to the next (resume) and is there for the back-trace
int32_t* cluster_cst(k2b_handle* h, int B, int K) {
  if (h->cg_next == nullptr) return nullptr;
  if (ensure(h, h->ws_cst, sizeof(int32_t) * (size_t)B * K) != K2B_OK) return nullptr;
  return static_cast<int32_t*>(h->ws_cst.p);
}

int32_t cluster_status(k2b_handle* h) {
  int st[4] = {0, 0, 0, 0};
  K2B_CUDA(h, cudaMemcpyAsync(st, h->dev_status, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  if (st[0] != 0 || st[1] != 0 || st[2] != 0) {
    cudaMemsetAsync(h->dev_status, 0, 3 * sizeof(int), h->stream);
    if (st[2] != 0 && st[0] == 0 && st[1] == 0)
      return fail(h, K2B_ERR_INVALID, "a Hyp passed by device pointer held a token id outside the vocabulary (it was decoded as blank)");
    return fail(h, K2B_ERR_STATE, st[0] ? "cluster search kernel: an mbarrier wait timed out" : "encoder_proj tcgen05 kernel: an mbarrier wait timed out");
  }
  return K2B_OK;
}

}  // namespace k2b
