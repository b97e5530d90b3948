// fp32 CUDA-core GEMM  C[M,N] = Aop[M,K] * W[N,K]^T (+ bias)  with the fused prologues/epilogues the
// transducer search needs (K2B_PREC_FP32 arithmetic).
//
//   PRO_DEC   A(m,k) = relu(tab0[y0(m)][k] + tab1[y1(m)][k])      stateless decoder: embedding gather +
//                                                                  grouped conv1d (folded) + ReLU, feeding
//                                                                  decoder_proj      (ref OfflineProjOfTransducer.cs:116)
//   PRO_JOIN  A(m,k) = tanh(enc[s(m)][k] + dec[m][k])              joiner prologue   (ref OfflineProjOfTransducer.cs:146)
//   EPI_TANH_ADD  C = tanh(acc + bias + enc[s(m)][n])              decoder_proj fused with the joiner prologue
//   EPI_EXP2X     C = exp(2*clamp(acc + bias, +-21))                  feeds the cluster kernel's tanh-from-exp prologue
//   EPI_ARGMAX    per (row, 64-column vocab tile) argmax, ties and NaN -> larger index
//                                                                  (ref OfflineRecognizer.cs:150-154)
//   EPI_TOPK      per (row, tile) max, sum-exp and top-k           (log_softmax + top-k of modified_beam_search)
//
// With EPI_ARGMAX / EPI_TOPK the logits never leave the SM: only a few words per (row, tile) are written.
#include <math.h>

#include "k2b_internal.h"

namespace k2b {

namespace {

constexpr int kThreads = 256;
constexpr int kPad = 4;

struct Best {
  float val;
  int idx;   // -1 = empty
  int nan;   // a NaN was met at or before idx inside the folded range
};

// Sequential-fold semantics of  token = logits[token] > logits[k] ? token : k  made associative:
// `a` covers lower indices than `b`.
__device__ __forceinline__ Best fold(const Best& a, const Best& b) {
  if (b.idx < 0) return a;
  if (a.idx < 0) return b;
  if (b.nan) return b;
  Best r = (a.val > b.val) ? a : b;
  r.nan = a.nan;
  return r;
}

__device__ __forceinline__ bool better(float v, int i, float ev, int ei) {
  return v > ev || (v == ev && i > ei);
}

template <int PRO>
__device__ __forceinline__ float4 load_a(const GemmArgs& a, bool row_ok, const float* p0, const float* p1, int k) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok) return v;
  if (PRO == PRO_PLAIN) {
    v = __ldg(reinterpret_cast<const float4*>(p0 + k));
  } else if (PRO == PRO_DEC) {
    float4 x = __ldg(reinterpret_cast<const float4*>(p0 + k));
    float4 y = __ldg(reinterpret_cast<const float4*>(p1 + k));
    v.x = fmaxf(x.x + y.x, 0.f);
    v.y = fmaxf(x.y + y.y, 0.f);
    v.z = fmaxf(x.z + y.z, 0.f);
    v.w = fmaxf(x.w + y.w, 0.f);
  } else {
    float4 x = __ldg(reinterpret_cast<const float4*>(p0 + k));
    float4 y = __ldg(reinterpret_cast<const float4*>(p1 + k));
    v.x = tanhf(x.x + y.x);
    v.y = tanhf(x.y + y.y);
    v.z = tanhf(x.z + y.z);
    v.w = tanhf(x.w + y.w);
  }
  return v;
}

__device__ __forceinline__ int table_row(int y, int V, int neg_wrap) {
  if (y >= 0) return y < V ? y : V;      // out-of-range ids read the zero row
  if (neg_wrap) { int r = y + V; return r >= 0 ? r : V; }
  return V;
}

template <int PRO, int EPI>
__global__ void __launch_bounds__(kThreads) gemm_simt_kernel(const GemmArgs a) {
  constexpr int kTileFloats = 2 * kBK * (kBM + kPad) + 2 * kBK * (kBN + kPad);
  constexpr int kEpiFloats = kBM * (kBN + 1);
  __shared__ __align__(16) float smem[kTileFloats > kEpiFloats ? kTileFloats : kEpiFloats];
  float (*As)[kBK][kBM + kPad] = reinterpret_cast<float (*)[kBK][kBM + kPad]>(smem);
  float (*Ws)[kBK][kBN + kPad] = reinterpret_cast<float (*)[kBK][kBN + kPad]>(smem + 2 * kBK * (kBM + kPad));

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int lr = tid >> 2, lk = (tid & 3) << 2;
  const int ty = tid >> 4, tx = tid & 15;

  // per-thread source rows of the loader
  const int am = m0 + lr;
  const bool a_ok = am < a.M;
  const float* p0 = nullptr;
  const float* p1 = nullptr;
  if (a_ok) {
    if (PRO == PRO_PLAIN) {
      p0 = a.A + (size_t)am * a.K;
    } else if (PRO == PRO_DEC) {
      int y0 = a.ctx[2 * am], y1 = a.ctx[2 * am + 1];
      if (a.compat_flag != nullptr && y0 < 0 && *a.compat_flag != 0) y0 = a.blank;
      p0 = a.tab0 + (size_t)table_row(y0, a.V, a.neg_wrap) * a.K;
      p1 = a.tab1 + (size_t)table_row(y1, a.V, a.neg_wrap) * a.K;
    } else {
      p0 = a.enc + (size_t)(am / a.rows_per_stream) * a.enc_stride;
      p1 = a.dec + (size_t)am * a.K;
    }
  }
  const int wn = n0 + lr;
  const bool w_ok = wn < a.N;
  const float* pw = a.W + (size_t)(w_ok ? wn : 0) * a.K;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra = load_a<PRO>(a, a_ok, p0, p1, lk);
  float4 rw = w_ok ? __ldg(reinterpret_cast<const float4*>(pw + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int nkb = a.K / kBK;
  int buf = 0;
  As[0][lk + 0][lr] = ra.x; As[0][lk + 1][lr] = ra.y; As[0][lk + 2][lr] = ra.z; As[0][lk + 3][lr] = ra.w;
  Ws[0][lk + 0][lr] = rw.x; Ws[0][lk + 1][lr] = rw.y; Ws[0][lk + 2][lr] = rw.z; Ws[0][lk + 3][lr] = rw.w;
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    if (kb + 1 < nkb) {
      const int k = (kb + 1) * kBK + lk;
      ra = load_a<PRO>(a, a_ok, p0, p1, k);
      rw = w_ok ? __ldg(reinterpret_cast<const float4*>(pw + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (kb + 1 < nkb) {
      const int nb = buf ^ 1;
      As[nb][lk + 0][lr] = ra.x; As[nb][lk + 1][lr] = ra.y; As[nb][lk + 2][lr] = ra.z; As[nb][lk + 3][lr] = ra.w;
      Ws[nb][lk + 0][lr] = rw.x; Ws[nb][lk + 1][lr] = rw.y; Ws[nb][lk + 2][lr] = rw.z; Ws[nb][lk + 3][lr] = rw.w;
      __syncthreads();
      buf = nb;
    }
  }

  // bias
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx * 4 + j;
    bv[j] = (a.bias != nullptr && n < a.N) ? __ldg(a.bias + n) : 0.f;
  }

  if (EPI == EPI_STORE || EPI == EPI_TANH_ADD || EPI == EPI_EXP2X) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m >= a.M) continue;
      const float* e = nullptr;
      if (EPI == EPI_TANH_ADD) e = a.enc + (size_t)(m / a.rows_per_stream) * a.enc_stride;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n >= a.N) continue;
        float v = acc[i][j] + bv[j];
        if (EPI == EPI_TANH_ADD) v = tanhf(v + __ldg(e + n));
        if (EPI == EPI_EXP2X) v = expf(2.f * fminf(fmaxf(v, -21.f), 21.f));
        a.C[(size_t)m * a.N + n] = v;
      }
    }
    return;
  }

  // ---- row-wise reductions over this CTA's 64-column vocab tile -------------------------------
  __syncthreads();  // everyone is done with As/Ws
  float (*Ct)[kBN + 1] = reinterpret_cast<float (*)[kBN + 1]>(smem);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Ct[ty * 4 + i][tx * 4 + j] = acc[i][j] + bv[j];
  __syncthreads();

  const int row = tid >> 2, sub = tid & 3;
  const int m = m0 + row;
  const int nvalid = min(kBN, a.N - n0);
  const int nt = gridDim.x;
  const int c_lo = sub * 16;
  const int c_hi = min(c_lo + 16, nvalid);

  if (EPI == EPI_ARGMAX) {
    Best b{0.f, -1, 0};
    for (int c = c_lo; c < c_hi; ++c) {
      const float v = Ct[row][c];
      Best e{v, n0 + c, (v != v) ? 1 : 0};
      b = fold(b, e);
    }
#pragma unroll
    for (int step = 1; step <= 2; step <<= 1) {
      Best o;
      o.val = __shfl_xor_sync(0xffffffffu, b.val, step);
      o.idx = __shfl_xor_sync(0xffffffffu, b.idx, step);
      o.nan = __shfl_xor_sync(0xffffffffu, b.nan, step);
      b = (sub & step) ? fold(o, b) : fold(b, o);
    }
    if (sub == 0 && m < a.M) {
      const size_t o = (size_t)m * nt + blockIdx.x;
      a.part_val[o] = b.val;
      a.part_idx[o] = b.idx;
      a.part_nan[o] = b.nan;
    }
    return;
  }

  if (EPI == EPI_TOPK) {
    const int K = a.topk;
    float tv[kMaxBeam];
    int ti[kMaxBeam];
#pragma unroll
    for (int i = 0; i < kMaxBeam; ++i) { tv[i] = -INFINITY; ti[i] = -1; }
    float mx = -INFINITY;
    for (int c = c_lo; c < c_hi; ++c) mx = fmaxf(mx, Ct[row][c]);
    float sum = 0.f;
    for (int c = c_lo; c < c_hi; ++c) {
      float v = Ct[row][c];
      sum += expf(v - mx);
      int idx = n0 + c;
#pragma unroll
      for (int i = 0; i < kMaxBeam; ++i) {
        if (i < K && better(v, idx, tv[i], ti[i])) {
          const float fv = tv[i]; const int fi = ti[i];
          tv[i] = v; ti[i] = idx; v = fv; idx = fi;
        }
      }
    }
#pragma unroll
    for (int step = 1; step <= 2; step <<= 1) {
      const float omx = __shfl_xor_sync(0xffffffffu, mx, step);
      const float osum = __shfl_xor_sync(0xffffffffu, sum, step);
      const float nm = fmaxf(mx, omx);
      if (nm == -INFINITY) {
        sum = 0.f;
      } else {
        sum = sum * expf(mx - nm) + osum * expf(omx - nm);
      }
      mx = nm;
      float pv[kMaxBeam];
      int pi[kMaxBeam];
#pragma unroll
      for (int i = 0; i < kMaxBeam; ++i) {
        pv[i] = __shfl_xor_sync(0xffffffffu, tv[i], step);
        pi[i] = __shfl_xor_sync(0xffffffffu, ti[i], step);
      }
#pragma unroll
      for (int q = 0; q < kMaxBeam; ++q) {
        if (q < K) {
          float v = pv[q];
          int idx = pi[q];
#pragma unroll
          for (int i = 0; i < kMaxBeam; ++i) {
            if (i < K && better(v, idx, tv[i], ti[i])) {
              const float fv = tv[i]; const int fi = ti[i];
              tv[i] = v; ti[i] = idx; v = fv; idx = fi;
            }
          }
        }
      }
    }
    if (sub == 0 && m < a.M) {
      const size_t o = (size_t)m * nt + blockIdx.x;
      a.part_m[o] = mx;
      a.part_s[o] = sum;
#pragma unroll
      for (int i = 0; i < kMaxBeam; ++i) {
        if (i < K) {
          a.part_tv[o * K + i] = tv[i];
          a.part_ti[o * K + i] = ti[i];
        }
      }
    }
  }
}

}  // namespace

int32_t launch_gemm_simt(k2b_handle* h, Pro pro, Epi epi, const GemmArgs& a) {
  if (a.M <= 0 || a.N <= 0) return K2B_OK;
  if (a.K <= 0 || (a.K % kBK) != 0) return fail(h, K2B_ERR_UNSUPPORTED, "GEMM inner dimension must be a multiple of 16");
  dim3 grid((a.N + kBN - 1) / kBN, (a.M + kBM - 1) / kBM);
  dim3 block(kThreads);
#define K2B_CASE(P, E)                                                         \
  if (pro == P && epi == E) {                                                  \
    gemm_simt_kernel<P, E><<<grid, block, 0, h->stream>>>(a);                  \
    K2B_LAUNCH_CHECK(h);                                                       \
    return K2B_OK;                                                             \
  }
  K2B_CASE(PRO_PLAIN, EPI_STORE)
  K2B_CASE(PRO_DEC, EPI_STORE)
  K2B_CASE(PRO_JOIN, EPI_STORE)
  K2B_CASE(PRO_DEC, EPI_TANH_ADD)
  K2B_CASE(PRO_PLAIN, EPI_ARGMAX)
  K2B_CASE(PRO_PLAIN, EPI_TOPK)
  K2B_CASE(PRO_PLAIN, EPI_EXP2X)
  K2B_CASE(PRO_DEC, EPI_EXP2X)
#undef K2B_CASE
  return fail(h, K2B_ERR_UNSUPPORTED, "GEMM prologue/epilogue combination not instantiated");
}

}  // namespace k2b
