// Greedy search of a FEW streams (cfg1: one utterance): one 8-CTA cluster per stream, the joiner output weight resident in
// REGISTERS, fp32 FMA - no tensor cores. At one stream the tcgen05 cluster kernel spends a frame on 64 MMAs for 32 hypothesis rows
// of which one is live (3.9 us per frame); here a frame is a 512 x 500 GEMV spread over 8 SMs:
//   * CTA `rank` owns 64 vocabulary rows, warp w four of them, lane l the joiner columns l, l + 32, .. (at most 16): 64 weight
//     registers per thread, loaded once per launch;
//   * per frame: x = tanh(enc[t] + dec(ctx)) from the two exp tables (one element per thread, the frame row prefetched two frames
//     ahead, the decoder row re-read only after an emission) -> shared memory -> 4 x 16 FMAs per lane -> butterfly -> the warp's
//     arg-best -> the CTA's -> one 16-byte st.async to each CTA of the cluster, completing on the destination's mbarrier ->
//     every CTA folds the eight candidates (deterministic: ties go to the larger index, ref OfflineRecognizer.cs:145-159) and
//     steps the context. No barrier.cluster in the loop.
// Replaces, for up to 16 streams of at least 16 frames, V <= 512, J <= 512: the beam-1 use of cluster_beam_kernel in
// k2b_greedy_offline (SINGLE / PER_STREAM, the two passes of BATCH_COMPAT) and k2b_greedy_online_chunk.
#include <limits.h>
#include <math.h>
#include <stdio.h>

#include "k2b_internal.h"
#include "sm100_ptx.cuh"

namespace k2b {

namespace {

using namespace ptx;

constexpr int kSgCS = 8;         // CTAs per cluster (= per stream)
constexpr int kSgWarps = 8;
constexpr int kSgRows = 8;       // vocabulary rows per warp: 128 weight registers per thread
constexpr int kSgKPL = 16;       // joiner columns per lane (J <= 512)
constexpr int kSgThreads = kSgWarps * 32;
constexpr int kSgXPT = 32 * kSgKPL / kSgThreads;    // joiner columns per thread in the prologue (2)
constexpr int kSgAhead = 3;      // frames of look-ahead of the frame-row loads

struct SgArgs {
  const float* encE;        // [B,Ttot,J]  exp(2*clamp(enc))
  const float* dec_tab;     // [(V+1)*V, J]  exp(2*clamp(decoder(y0,y1)))
  const float* out_w;       // [V,J]
  const float* out_b;       // [V]
  int B, T, V, J, blank, unk, extra_mask, t0, Ttot, cap;
  const int64_t* hyp_in;    // [B,2] or null = {-1, blank}
  int64_t* hyp_out;         // [B,2] or null
  const int32_t* lens;      // [B] or null
  int64_t* tokens;          // [B,cap]
  int32_t* ts;              // [B,cap]
  int32_t* n_out;           // [B]
  int* status;
  int debug;
};

__device__ __forceinline__ bool sg_better(float v, int i, float ev, int ei) { return v > ev || (v == ev && i > ei); }
// order-preserving integer image of a logit (NaN and "no candidate" below everything): the warp's arg-best is then two REDUX
// (the maximum key, then the largest index among the lanes that hold it - ties go to the larger index, ref OfflineRecognizer.cs:150-154)
constexpr int kSgNone = (int)0x80000000;
__device__ __forceinline__ int sg_key(float v, int i) {
  const int k = __float_as_int(v);
  return (i >= 0 && v == v) ? (k ^ ((k >> 31) & 0x7fffffff)) : kSgNone;
}
__device__ __forceinline__ int sg_warp_best(float v, int i) {     // returns the winning index (-1: no candidate in the warp)
  const int key = sg_key(v, i);
  const int wk = __reduce_max_sync(0xffffffffu, key);
  return __reduce_max_sync(0xffffffffu, (key == wk && key != kSgNone) ? i : -1);
}

// tanh(e + d) from exp(2e), exp(2d) (both clamped when the tables were built): one MUFU
__device__ __forceinline__ float sg_tanh(float ee, float ed) {
  const float y = fmaf(ee, ed, 1.0f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return fmaf(-2.0f, r, 1.0f);
}

// latency-critical wait of a few lanes: mbarrier.test_wait in a tight loop (the hardware-suspended try_wait of sm100_ptx.cuh took
// ~1 k cycles to come back here, 2.2 k when only eight lanes waited)
__device__ __forceinline__ bool sg_wait_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int spins = 0; !done; ++spins) {
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && (spins & 1023) == 1023 && clock64() - t0 > kWaitTimeoutCycles) return false;
  }
  return true;
}

// a.debug: thread 0 of CTA 0 prints phase cycle totals (tools/time_single_greedy.py)
__global__ void __cluster_dims__(kSgCS, 1, 1) __launch_bounds__(kSgThreads, 1) single_greedy_kernel(const SgArgs a) {
  __shared__ __align__(16) float xs[32 * kSgKPL];
  __shared__ __align__(16) uint4 mail[2][kSgCS * kSgWarps];   // [frame parity][source CTA, warp]: (logit bits, vocabulary index, -, -)
  __shared__ uint64_t mb[2];
  __shared__ int ysh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.x / kSgCS;
  const int J = a.J, V = a.V, T = a.T;
  if (tid == 0) { mbar_init(&mb[0], 1); mbar_init(&mb[1], 1); mbar_fence_init(); }
#pragma unroll
  for (int u = 0; u < kSgXPT; ++u) xs[tid + u * kSgThreads] = 0.f;       // columns beyond J stay zero (zero weights against them)

  // ---- this thread's part of the weight: rows row0 .. row0 + 7, columns lane + 32 i ----------------------------------------
  const int row0 = (int)rank * (kSgWarps * kSgRows) + warp * kSgRows;
  float2 w2[kSgRows / 2][kSgKPL];       // rows in pairs: one FFMA2 (two fp32 FMAs, each rounded like fmaf) per pair and column
  float bias[kSgRows];
#pragma unroll
  for (int r = 0; r < kSgRows; ++r) {
    const int row = row0 + r;
#pragma unroll
    for (int i = 0; i < kSgKPL; ++i) {
      const int k = lane + 32 * i;
      const float wv = (row < V && k < J) ? __ldg(a.out_w + (size_t)row * J + k) : 0.f;
      if (r & 1) w2[r >> 1][i].y = wv; else w2[r >> 1][i].x = wv;
    }
    bias[r] = row < V ? __ldg(a.out_b + row) : -INFINITY;
  }
  int c0 = -1, c1 = a.blank;
  if (a.hyp_in != nullptr) { c0 = (int)a.hyp_in[2 * b]; c1 = (int)a.hyp_in[2 * b + 1]; }
  int n = 0;
  const int len_b = a.lens != nullptr ? a.lens[b] : INT_MAX;
  const float* erow = a.encE + ((size_t)b * a.Ttot + a.t0) * J;
  // prologue: this thread's joiner columns tid, tid + 256; their frame values kSgAhead frames ahead, the decoder row re-read on emission
  float e[kSgAhead][kSgXPT], d[kSgXPT];
#pragma unroll
  for (int u = 0; u < kSgXPT; ++u) {
    const int k = tid + u * kSgThreads;
#pragma unroll
    for (int f = 0; f < kSgAhead; ++f) e[f][u] = (k < J && f < T) ? __ldg(erow + (size_t)f * J + k) : 0.f;
    d[u] = k < J ? __ldg(a.dec_tab + ((size_t)(c0 + 1) * V + c1) * J + k) : 0.f;
  }
  bool ok = true;
  const bool dbg = a.debug != 0 && blockIdx.x == 0 && tid == 0;
  long long tph[5] = {0, 0, 0, 0, 0}, tl = 0;
  __syncthreads();
  cluster_sync();                                  // every CTA's mbarriers exist before the first remote store
  if (dbg) tl = clock64();

  for (int t = 0; t < T; ++t) {
    const int par = t & 1;
#pragma unroll
    for (int u = 0; u < kSgXPT; ++u) {
      const int k = tid + u * kSgThreads;
      if (k < J) xs[k] = sg_tanh(e[0][u], d[u]);
#pragma unroll
      for (int f = 0; f + 1 < kSgAhead; ++f) e[f][u] = e[f + 1][u];
      if (k < J && t + kSgAhead < T) e[kSgAhead - 1][u] = __ldg(erow + (size_t)(t + kSgAhead) * J + k);
    }
    if (tid == 0) mbar_expect_tx(&mb[par], (uint32_t)(kSgCS * kSgWarps * 16));
    __syncthreads();
    if (dbg) { const long long now = clock64(); tph[0] += now - tl; tl = now; }
    // ---- eight dot products per warp ------------------------------------------------------------------------------------
    float2 acc2[kSgRows / 2];
#pragma unroll
    for (int r = 0; r < kSgRows / 2; ++r) acc2[r] = make_float2(0.f, 0.f);
    float xv[kSgKPL];
#pragma unroll
    for (int i = 0; i < kSgKPL; ++i) xv[i] = xs[lane + 32 * i];
#pragma unroll
    for (int i = 0; i < kSgKPL; ++i) {
      const float2 x2 = make_float2(xv[i], xv[i]);
#pragma unroll
      for (int r = 0; r < kSgRows / 2; ++r) acc2[r] = __ffma2_rn(w2[r][i], x2, acc2[r]);
    }
    float acc[kSgRows];
#pragma unroll
    for (int r = 0; r < kSgRows / 2; ++r) { acc[2 * r] = acc2[r].x; acc[2 * r + 1] = acc2[r].y; }
    // transposed butterfly: after the steps over lane bits 4, 3, 2 a lane holds ONE row's partial sum (row = bits 4..2 of its
    // id, most significant first), then two plain steps finish it: 4 + 2 + 1 + 1 + 1 shuffles instead of 5 x 8
    float a4[4], a2[2], a1;
    {
      const bool up = (lane & 16) != 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float send = up ? acc[r] : acc[r + 4], keep = up ? acc[r + 4] : acc[r];
        a4[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    {
      const bool up = (lane & 8) != 0;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float send = up ? a4[r] : a4[r + 2], keep = up ? a4[r + 2] : a4[r];
        a2[r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    }
    {
      const bool up = (lane & 4) != 0;
      const float send = up ? a2[0] : a2[1], keep = up ? a2[1] : a2[0];
      a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    const int myr = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);      // the row this lane ended up with
    float bv = -INFINITY;
    int bi = -1;
    {
      float bsel = bias[0];
#pragma unroll
      for (int r = 1; r < kSgRows; ++r) bsel = myr == r ? bias[r] : bsel;
      const int idx = row0 + myr;
      if (idx < V) { bv = a1 + bsel; bi = idx; }
    }
    {                                              // the warp's best over its eight rows: its value travels with it
      const int wi_ = sg_warp_best(bv, bi);
      const float wv_ = __shfl_sync(0xffffffffu, bv, wi_ >= 0 ? ((wi_ - row0) << 2) : 0);     // lanes 4 r .. 4 r + 3 hold row r
      bv = wi_ >= 0 ? wv_ : -INFINITY;
      bi = wi_;
    }
    if (dbg) { const long long now = clock64(); tph[1] += now - tl; tl = now; }
    if (lane < kSgCS)         // lane q delivers this warp's candidate to CTA q (its own mailbox included)
      dsmem_st_async_v4(dsmem_map(smem_u32(&mail[par][rank * kSgWarps + warp]), (uint32_t)lane), __float_as_uint(bv), (uint32_t)bi, 0u, 0u,
                        dsmem_map(smem_u32(&mb[par]), (uint32_t)lane));
    // ---- the frame's symbol: warp 0 of every CTA folds the same 64 candidates, the other warps pick the result up behind the
    //      CTA barrier
    if (warp == 0) {
      if (!sg_wait_spin(&mb[par], (uint32_t)((t >> 1) & 1))) ok = false;
      if (dbg) { const long long now = clock64(); tph[2] += now - tl; tl = now; }
      float v = -INFINITY;
      int i = -1;
#pragma unroll
      for (int u = 0; u < kSgCS * kSgWarps / 32; ++u) {
        const uint4 m = mail[par][lane + 32 * u];
        const float mv = __uint_as_float(m.x);
        const int mi = (int)m.y;
        if (mi >= 0 && (i < 0 || sg_better(mv, mi, v, i))) { v = mv; i = mi; }
      }
      const int yb = sg_warp_best(v, i);
      if (lane == 0) ysh = yb;
    }
    __syncthreads();
    const int y = ysh;
    if (dbg) { const long long now = clock64(); tph[3] += now - tl; tl = now; }
    const bool emit = y >= 0 && y != a.blank && y != a.unk && y != a.extra_mask && a.t0 + t < len_b;
    if (emit) {
      if (rank == 0 && tid == 0 && n < a.cap) {
        a.tokens[(size_t)b * a.cap + n] = y;
        a.ts[(size_t)b * a.cap + n] = a.t0 + t;
      }
      ++n;
      c0 = c1;
      c1 = y;
#pragma unroll
      for (int u = 0; u < kSgXPT; ++u) {
        const int k = tid + u * kSgThreads;
        if (k < J) d[u] = __ldg(a.dec_tab + ((size_t)(c0 + 1) * V + c1) * J + k);
      }
    }
    if (dbg) { const long long now = clock64(); tph[4] += now - tl; tl = now; }
  }
  if (dbg)
    printf("single_greedy phases (cycles per frame, CTA 0 thread 0): prologue+sync %lld | dot+butterfly+warp best %lld | send+wait %lld | "
           "fold+sync %lld | emit %lld\n", tph[0] / T, tph[1] / T, tph[2] / T, tph[3] / T, tph[4] / T);
  if (rank == 0 && tid == 0) {
    a.n_out[b] = n;
    if (a.hyp_out != nullptr) { a.hyp_out[2 * b] = c0; a.hyp_out[2 * b + 1] = c1; }
  }
  if (!ok) atomicExch(a.status, 1);
  cluster_sync();                                  // nobody leaves while a peer may still store into its mailbox
}

}  // namespace

bool single_greedy_usable(const k2b_handle* h, int B, int T) {
  const k2b_config& c = h->cfg;
  return h->opt_single_greedy != 0 && B >= 1 && B <= 16 && T >= 16 && c.vocab_size <= kSgCS * kSgWarps * kSgRows &&
         c.joiner_dim <= 32 * kSgKPL && c.joiner_dim % 32 == 0 && h->dec_tab != nullptr;
}

// encE: [B,Ttot,J] frames mapped through exp(2x) (as the cluster kernel takes them); decodes frames [t0, t0 + T)
int32_t single_greedy_dev(k2b_handle* h, const float* encE, int B, int T, int t0, int Ttot, int extra_mask, const int64_t* hyp_in,
                          int64_t* hyp_out, int64_t* tokens, int32_t* ts, int32_t* n_out, int cap) {
  const k2b_config& c = h->cfg;
  SgArgs a;
  a.encE = encE; a.dec_tab = h->dec_tab; a.out_w = h->out_w; a.out_b = h->out_b;
  a.B = B; a.T = T; a.V = c.vocab_size; a.J = c.joiner_dim; a.blank = c.blank_id; a.unk = c.unk_id; a.extra_mask = extra_mask;
  a.t0 = t0; a.Ttot = Ttot; a.cap = cap;
  a.hyp_in = hyp_in; a.hyp_out = hyp_out; a.lens = h->lens_active ? h->lens_dev : nullptr;
  a.tokens = tokens; a.ts = ts; a.n_out = n_out; a.status = h->dev_status;
  a.debug = h->opt_single_greedy >= 2 ? 1 : 0;
  prof_begin(h);
  single_greedy_kernel<<<B * kSgCS, kSgThreads, 0, h->stream>>>(a);
  prof_end(h);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

}  // namespace k2b
