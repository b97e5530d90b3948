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
//   * epilogue: TMEM -> registers (+bias) -> shared transpose -> per hypothesis max / sum-exp / top-K with
//     warp REDUX; the (2+2K)-word partial of every (slice, hypothesis) is written into every CTA of the cluster
//     through distributed shared memory; one cluster barrier; every CTA then runs the hypothesis merge
//     (log_softmax constants, per-stream top-K, dedupe by sequence hash, log-add, prune) redundantly and
//     deterministically, so no second exchange is needed. CTA 0 records the back-pointers.
#include <math.h>

#include "k2b_internal.h"
#include "sm100_ptx.cuh"

namespace k2b {

namespace {

using namespace ptx;

constexpr int kNH = 32;          // hypothesis rows per cluster (UMMA N)
constexpr int kCThreads = 512;
constexpr int kLtStride = 33;
constexpr uint64_t kHashSeedC = 0x9E3779B97F4A7C15ull;

struct HypState {
  int ctx0[kNH], ctx1[kNH];
  float lp[kNH];
  int len[kNH];
  unsigned long long hash[kNH];
  int nlive[kNH];
};

struct ClusterArgs {
  const float* encE;        // [B,T,J]  exp(2*clamp(enc))
  const float* dec_tab;     // [(V+1)*V, J]  exp(2*clamp(decoder(y0,y1)))
  const uint8_t* wo_hi_img; // [CS][J/64][128 x 128 B swizzled]
  const uint32_t* wo_lo;    // [CS*128][J/2] packed bf16 pairs
  const float* bias;        // [CS*128], -inf beyond V
  int B, T, K, V, J, S, CS, blank, unk, x3;
  int32_t* bp;              // [B,T,K]
  float* fin_lp;            // [B*K]
  int32_t* fin_len;         // [B*K]
  int32_t* fin_nlive;       // [B]
  int* status;
};

__device__ __forceinline__ int fkey(float f) { const int k = __float_as_int(f); return k >= 0 ? k : (k ^ 0x7fffffff); }
__device__ __forceinline__ float funkey(int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff)); }
__device__ __forceinline__ bool better_c(float v, int i, float ev, int ei) { return v > ev || (v == ev && i > ei); }

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
  return mx + log1pf(expf(mn - mx));
}

// tanh(e + d) from Ee = exp(2e), Ed = exp(2d)
__device__ __forceinline__ float tanh_from_exp(float ee, float ed) {
  const float y = fmaf(ee, ed, 1.0f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  r = r * fmaf(-y, r, 2.0f);                 // one Newton step: full fp32 accuracy, inf -> handled below
  const float x = fmaf(-2.0f, r, 1.0f);
  return (y > 3.0e38f) ? 1.0f : x;           // y = inf: r = 0 * (2 - inf*0 = NaN) -> select the limit
}

// One warp: hypothesis merge of local stream s (see beam_select_kernel in search.cu for the global-memory twin).
__device__ void select_stream(int s, int K, int V, int CS, int XW, const float* __restrict__ xb, const HypState& in,
                              HypState& out, int blank, int unk, int32_t* __restrict__ bp_row, int lane) {
  const unsigned full = 0xffffffffu;
  const int nl = in.nlive[s];
  if (nl == 0) {
    if (lane < K) {
      const int o = s * K + lane;
      out.ctx0[o] = -1; out.ctx1[o] = blank; out.lp[o] = -INFINITY; out.len[o] = 2; out.hash[o] = kHashSeedC;
    }
    if (lane == 0) out.nlive[s] = 0;
    return;
  }
  float M = -INFINITY, L = 0.f, LP = 0.f;
  if (lane < nl) {
    const int n = s * K + lane;
    for (int c = 0; c < CS; ++c) M = fmaxf(M, xb[(c * kNH + n) * XW]);
    float sum = 0.f;
    for (int c = 0; c < CS; ++c) {
      const float pm = xb[(c * kNH + n) * XW], ps = xb[(c * kNH + n) * XW + 1];
      if (pm != -INFINITY) sum += ps * expf(pm - M);
    }
    L = logf(sum);
    LP = in.lp[n];
  }
  float tv[kMaxBeam];
  int tf[kMaxBeam];
#pragma unroll
  for (int i = 0; i < kMaxBeam; ++i) { tv[i] = -INFINITY; tf[i] = -1; }
  const int per_h = CS * K, total = nl * per_h;
  for (int base = 0; base < total; base += 32) {
    const int c = base + lane;
    const bool valid = c < total;
    const int h = valid ? c / per_h : 0;
    const float Mh = __shfl_sync(full, M, h), Lh = __shfl_sync(full, L, h), LPh = __shfl_sync(full, LP, h);
    if (valid) {
      const int r = c - h * per_h, slice = r / K, j = r - slice * K;
      const float* e = xb + (slice * kNH + s * K + h) * XW;
      const int idx = __float_as_int(e[2 + K + j]);
      if (idx >= 0 && idx < V) {
        float v = ((e[2 + j] - Mh) - Lh) + LPh;   // same operation order as log_softmax(x) + lp
        int f = h * V + idx;
        if (v == v) {
#pragma unroll
          for (int i = 0; i < kMaxBeam; ++i) {
            if (i < K && better_c(v, f, tv[i], tf[i])) {
              const float fv = tv[i]; const int ff = tf[i];
              tv[i] = v; tf[i] = f; v = fv; f = ff;
            }
          }
        }
      }
    }
  }
  float my_v = -INFINITY;
  int my_f = -1;
#pragma unroll
  for (int r = 0; r < kMaxBeam; ++r) {
    if (r < K) {
      const int hk = tf[0] >= 0 ? fkey(tv[0]) : (int)0x80000000;
      const int wk = __reduce_max_sync(full, hk);
      const int cf = (tf[0] >= 0 && hk == wk) ? tf[0] : -1;
      const int wf = __reduce_max_sync(full, cf);
      if (wf >= 0 && tf[0] == wf) {
#pragma unroll
        for (int i = 0; i + 1 < kMaxBeam; ++i) { tv[i] = tv[i + 1]; tf[i] = tf[i + 1]; }
        tv[kMaxBeam - 1] = -INFINITY; tf[kMaxBeam - 1] = -1;
      }
      if (lane == r) { my_v = funkey(wk); my_f = wf; }
    }
  }
  const bool cand = lane < K && my_f >= 0;
  int par = 0, tok = -1, c0 = -1, c1 = blank, ln = 2;
  uint64_t hs = kHashSeedC;
  if (cand) {
    par = my_f / V;
    const int y = my_f - par * V;
    const int prow = s * K + par;
    hs = in.hash[prow]; ln = in.len[prow]; c0 = in.ctx0[prow]; c1 = in.ctx1[prow];
    if (y != blank && y != unk) { tok = y; hs = hash_push_c(hs, y); ln += 1; c0 = c1; c1 = y; }
  }
  int root = lane;
  for (int q = 0; q < K; ++q) {
    const uint64_t qh = __shfl_sync(full, hs, q);
    const int ql = __shfl_sync(full, ln, q), q0 = __shfl_sync(full, c0, q), q1 = __shfl_sync(full, c1, q);
    const int qc = __shfl_sync(full, (int)cand, q);
    if (cand && qc && q < lane && root == lane && qh == hs && ql == ln && q0 == c0 && q1 == c1) root = q;
  }
  float lp = my_v;
  for (int q = 0; q < K; ++q) {
    const int qroot = __shfl_sync(full, root, q);
    const float qv = __shfl_sync(full, my_v, q);
    const int qc = __shfl_sync(full, (int)cand, q);
    if (cand && qc && q != lane && qroot == lane) lp = logaddexp_c(lp, qv);
  }
  const bool is_root = cand && root == lane;
  const unsigned roots = __ballot_sync(full, is_root);
  const int nnew = __popc(roots);
  if (is_root) {
    const int slot = __popc(roots & ((1u << lane) - 1u));
    const int o = s * K + slot;
    out.ctx0[o] = c0; out.ctx1[o] = c1; out.lp[o] = lp; out.len[o] = ln; out.hash[o] = hs;
    if (bp_row != nullptr) bp_row[slot] = (par << 28) | (tok + 1);
  }
  if (lane >= nnew && lane < K) {
    const int o = s * K + lane;
    out.ctx0[o] = -1; out.ctx1[o] = blank; out.lp[o] = -INFINITY; out.len[o] = 2; out.hash[o] = kHashSeedC;
    if (bp_row != nullptr) bp_row[lane] = 0;
  }
  if (lane == 0) out.nlive[s] = nnew;
}

__global__ void __launch_bounds__(kCThreads, 1) cluster_beam_kernel(const ClusterArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ HypState st[2];
  __shared__ float bias_s[128];
  __shared__ uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x / a.CS;
  const int J = a.J, K = a.K, V = a.V, S = a.S, CS = a.CS, T = a.T;
  const int nkb = J / 64;
  const int XW = 2 + 2 * K;
  uint8_t* w_hi = smem;
  uint8_t* b_hi = w_hi + (size_t)nkb * 16384;
  uint8_t* b_lo = b_hi + (size_t)nkb * kNH * 128;
  float* Lt = reinterpret_cast<float*>(b_hi);                      // aliases the B operand between MMA and next build
  float* xch = reinterpret_cast<float*>(b_lo + (size_t)nkb * kNH * 128);   // [2][CS][NH][XW]

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t t_d = tbase, t_wlo = tbase + 64;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;

  // ---- one-time staging of this CTA's weight slice ------------------------------------------------------
  if (tid == 0) {
    mbar_expect_tx(&bar_w, (uint32_t)(nkb * 16384));
    for (int kb = 0; kb < nkb; ++kb)
      tma_bulk_g2s(w_hi + (size_t)kb * 16384, a.wo_hi_img + ((size_t)rank * nkb + kb) * 16384, 16384, &bar_w);
  }
  if (a.x3 && warp < 4) {
    const uint32_t* src = a.wo_lo + ((size_t)rank * 128 + tid) * (J / 2);
    for (int c0 = 0; c0 < J / 2; c0 += 32) {
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
    const bool used = n < S * K;
    for (int b = 0; b < 2; ++b) {
      st[b].ctx0[n] = -1; st[b].ctx1[n] = a.blank;
      st[b].lp[n] = (used && h == 0) ? 0.f : -INFINITY;
      st[b].len[n] = 2; st[b].hash[n] = kHashSeedC;
      st[b].nlive[n] = 0;
    }
    if (n < S) st[0].nlive[n] = (cluster * S + n < a.B) ? 1 : 0;
  }
  bool ok = true;
  if (tid == 0) ok = mbar_wait(&bar_w, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync();            // every CTA of the cluster is resident: remote shared memory may be written

  const uint32_t idesc = umma_idesc_bf16_f32(128, kNH);
  const uint32_t w_hi_u = smem_u32(w_hi), b_hi_u = smem_u32(b_hi), b_lo_u = smem_u32(b_lo);
  const int nq = J / 4;
  int cur = 0;

  for (int t = 0; t < T; ++t) {
    // ---- (a) joiner prologue: x[n,:] = tanh(enc[stream(n),t,:] + dec(ctx(n))) as bf16 hi/lo, K-major swizzled --
    {
      const HypState& sc = st[cur];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int n = warp * 2 + r;
        int s = n / K;
        if (s >= S) s = S - 1;
        int g = cluster * S + s;
        if (g >= a.B) g = a.B - 1;
        const float4* pe = reinterpret_cast<const float4*>(a.encE + ((size_t)g * T + t) * J);
        const float4* pd = reinterpret_cast<const float4*>(a.dec_tab + ((size_t)(sc.ctx0[n] + 1) * V + sc.ctx1[n]) * J);
        float4 e[4], d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int q = lane + 32 * i;
          if (q < nq) { e[i] = __ldg(pe + q); d[i] = __ldg(pd + q); }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int q = lane + 32 * i;
          if (q < nq) {
            const float x0 = tanh_from_exp(e[i].x, d[i].x), x1 = tanh_from_exp(e[i].y, d[i].y);
            const float x2 = tanh_from_exp(e[i].z, d[i].z), x3 = tanh_from_exp(e[i].w, d[i].w);
            const int k = 4 * q;
            const uint32_t off = (uint32_t)(k >> 6) * (kNH * 128) + sw128_offset(n, k & 63);
            const float h0 = bf16_round(x0), h1 = bf16_round(x1), h2 = bf16_round(x2), h3 = bf16_round(x3);
            *reinterpret_cast<uint2*>(b_hi + off) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
            if (a.x3)
              *reinterpret_cast<uint2*>(b_lo + off) = make_uint2(pack_bf16x2(x0 - h0, x1 - h1), pack_bf16x2(x2 - h2, x3 - h3));
          }
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- (b) D[128 vocab, 32 hyps] = W_slice * x^T on the tensor core (one thread issues) --------------------
    if (tid == 0) {
      uint32_t acc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dw = umma_desc_k_sw128(w_hi_u + kb * 16384 + k * 32);
          const uint64_t dxh = umma_desc_k_sw128(b_hi_u + kb * (kNH * 128) + k * 32);
          umma_ss(t_d, dw, dxh, idesc, acc);
          acc = 1;
          if (a.x3) {
            const uint64_t dxl = umma_desc_k_sw128(b_lo_u + kb * (kNH * 128) + k * 32);
            umma_ss(t_d, dw, dxl, idesc, 1);
            umma_ts(t_d, t_wlo + (uint32_t)((kb * 4 + k) * 8), dxh, idesc, 1);
          }
        }
      }
      umma_commit(&bar_mma);
    }
    if (!mbar_wait(&bar_mma, (uint32_t)(t & 1))) ok = false;
    tc_fence_after();

    // ---- (c) accumulator -> registers (+bias) -> transposed shared tile ------------------------------------------
    if (warp < 4) {
      uint32_t v[32];
      tmem_ld32(t_d + lane_base, v);
      tmem_ld_wait();
      const float bsv = bias_s[tid];
#pragma unroll
      for (int n = 0; n < kNH; ++n) Lt[tid * kLtStride + n] = __uint_as_float(v[n]) + bsv;
    }
    tc_fence_before();
    __syncthreads();

    // ---- (d) per hypothesis: max, sum-exp and top-K over this slice's 128 logits (warp REDUX) ----------------------
    float* xw = xch + (size_t)(t & 1) * CS * kNH * XW;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int n = warp * 2 + r;
      float v[4];
      int key[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { v[j] = Lt[(lane + 32 * j) * kLtStride + n]; key[j] = fkey(v[j]); }
      const int kmax = max(max(key[0], key[1]), max(key[2], key[3]));
      const float m = funkey(__reduce_max_sync(0xffffffffu, kmax));
      float sum = 0.f;
      if (m != -INFINITY) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sum += expf(v[j] - m);
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      unsigned taken = 0;
      float out_v = -INFINITY;
      int out_i = -1;
      for (int rr = 0; rr < K; ++rr) {
        int bk = (int)0x80000000, bj = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (!((taken >> j) & 1u) && key[j] >= bk) { bk = key[j]; bj = j; }
        const int wk = __reduce_max_sync(0xffffffffu, bk);
        const int ci = (bj >= 0 && bk == wk) ? (int)(rank * 128 + lane + 32 * bj) : -1;
        const int wi = __reduce_max_sync(0xffffffffu, ci);
        if (ci == wi && wi >= 0) taken |= 1u << bj;
        if (lane == rr) { out_v = funkey(wk); out_i = (wi >= 0 && wi < V) ? wi : -1; }
      }
      // partial of (slice = rank, hypothesis n) -> every CTA of the cluster
      const uint32_t base = smem_u32(xw + ((size_t)rank * kNH + n) * XW);
      for (uint32_t dst = 0; dst < (uint32_t)CS; ++dst) {
        const uint32_t rb = dsmem_map(base, dst);
        if (lane < K) {
          dsmem_st_f32(rb + 4u * (2 + lane), out_v);
          dsmem_st_u32(rb + 4u * (2 + K + lane), (uint32_t)out_i);
        } else if (lane == K) {
          dsmem_st_f32(rb, m);
        } else if (lane == K + 1) {
          dsmem_st_f32(rb + 4u, sum);
        }
      }
    }
    cluster_arrive();
    cluster_wait();

    // ---- (e) hypothesis merge, redundantly in every CTA --------------------------------------------------------
    for (int s = warp; s < S; s += kCThreads / 32) {
      const int g = cluster * S + s;
      int32_t* bp_row = (rank == 0 && g < a.B) ? a.bp + ((size_t)g * T + t) * K : nullptr;
      select_stream(s, K, V, CS, XW, xw, st[cur], st[cur ^ 1], a.blank, a.unk, bp_row, lane);
    }
    __syncthreads();
    cur ^= 1;
  }

  if (rank == 0 && tid < S * K) {
    const int s = tid / K, hslot = tid % K, g = cluster * S + s;
    if (g < a.B) {
      a.fin_lp[(size_t)g * K + hslot] = st[cur].lp[tid];
      a.fin_len[(size_t)g * K + hslot] = st[cur].len[tid];
      if (hslot == 0) a.fin_nlive[g] = st[cur].nlive[s];
    }
  }
  if (!ok) atomicExch(a.status, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
  cluster_sync();
}

// ---- weight-load-time packing ----------------------------------------------------------------------------------
__global__ void pack_out_w_kernel(const float* __restrict__ out_w, const float* __restrict__ out_b, int V, int J, int CS,
                                  uint8_t* __restrict__ img, uint32_t* __restrict__ lo, float* __restrict__ bias) {
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
  if (i < n) out[i] = expf(2.f * fminf(fmaxf(in[i], -40.f), 40.f));
}

}  // namespace

bool cluster_path_supported(const k2b_handle* h, int K) {
  const k2b_config& c = h->cfg;
  const int CS = (c.vocab_size + 127) / 128;
  if (c.vocab_size > 1024 || c.joiner_dim > 512 || c.joiner_dim % 64) return false;
  if (K < 1 || K > kMaxBeam) return false;
  const size_t dyn = (size_t)(c.joiner_dim / 64) * (16384 + 2 * kNH * 128) + 2ull * CS * kNH * (2 + 2 * K) * 4;
  if (dyn + 4096 > 227 * 1024) return false;
  const size_t tab = (size_t)(c.vocab_size + 1) * c.vocab_size * c.joiner_dim * sizeof(float);
  return tab <= ((size_t)16 << 30);
}

// Builds (once per weight load) the shared-memory image / TMEM source / padded bias of the joiner weight and the
// memoised decoder table.
int32_t ensure_cluster_assets(k2b_handle* h) {
  if (h->tc_ready) return K2B_OK;
  const k2b_config& c = h->cfg;
  const int V = c.vocab_size, J = c.joiner_dim, D = c.decoder_dim, CS = (V + 127) / 128;
  const size_t rows = (size_t)CS * 128;
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->wo_hi_img), rows * J * 2));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->wo_lo), rows * J * 2));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->bias_pad), rows * sizeof(float)));
  pack_out_w_kernel<<<(unsigned)rows, 128, 0, h->stream>>>(h->out_w, h->out_b, V, J, CS, h->wo_hi_img, h->wo_lo, h->bias_pad);
  K2B_LAUNCH_CHECK(h);
  const long long nctx = (long long)(V + 1) * V;
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->dec_tab), (size_t)nctx * J * sizeof(float)));
  const int chunk = 1 << 18;
  int32_t* ctx = nullptr;
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&ctx), sizeof(int32_t) * 2 * chunk));
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
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(ctx);
  h->tc_ready = true;
  return K2B_OK;
}

int32_t exp2x_frames(k2b_handle* h, const float* in, float* out, size_t n) {
  exp2x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(in, out, n);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

// encE: [B,T,J] frames already mapped through exp(2x). Writes bp + final state; the caller runs the back-trace.
int32_t beam_cluster_dev(k2b_handle* h, const float* encE, int B, int T, int K, int32_t* bp, float* fin_lp, int32_t* fin_len,
                         int32_t* fin_nlive) {
  const k2b_config& c = h->cfg;
  const int V = c.vocab_size, J = c.joiner_dim, CS = (V + 127) / 128;
  const int S = kNH / K;
  const int nclusters = (B + S - 1) / S;
  K2B_TRY(ensure(h, h->ws_misc, 256));
  int* status = static_cast<int*>(h->ws_misc.p);
  K2B_CUDA(h, cudaMemsetAsync(status, 0, sizeof(int), h->stream));
  ClusterArgs a;
  a.encE = encE; a.dec_tab = h->dec_tab; a.wo_hi_img = h->wo_hi_img; a.wo_lo = h->wo_lo; a.bias = h->bias_pad;
  a.B = B; a.T = T; a.K = K; a.V = V; a.J = J; a.S = S; a.CS = CS; a.blank = c.blank_id; a.unk = c.unk_id;
  a.x3 = c.precision == K2B_PREC_BF16X3 ? 1 : 0;
  a.bp = bp; a.fin_lp = fin_lp; a.fin_len = fin_len; a.fin_nlive = fin_nlive; a.status = status;
  const size_t dyn = (size_t)(J / 64) * (16384 + 2 * kNH * 128) + 2ull * CS * kNH * (2 + 2 * K) * 4;
  K2B_CUDA(h, cudaFuncSetAttribute(cluster_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nclusters * CS));
  cfg.blockDim = dim3(kCThreads);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  prof_begin(h);
  K2B_CUDA(h, cudaLaunchKernelEx(&cfg, cluster_beam_kernel, a));
  prof_end(h);
  h->launches++;
  return K2B_OK;
}

int32_t cluster_status(k2b_handle* h) {
  int st = 0;
  K2B_CUDA(h, cudaMemcpyAsync(&st, h->ws_misc.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  if (st != 0) return fail(h, K2B_ERR_STATE, "cluster search kernel: an mbarrier wait timed out");
  return K2B_OK;
}

}  // namespace k2b
