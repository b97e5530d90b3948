// Fused CTC greedy search: per-frame argmax -> blank/repeat collapse -> trailing-blank count, one launch.
//
// Restates ForwardGreedySearchCTC / ForwardBatchGreedySearchCTC (ref OfflineRecognizer.cs:305-430) and the
// online variant (ref OnlineRecognizer.cs:220-319):
//   y = first index of the frame maximum            (ref :335  IndexOf(Max), ties -> lowest index, Q2)
//   trailing = (y == blank) ? trailing + 1 : 0      (ref :337-344)
//   emit y at t + frame_offset when y != blank && y != prev_id; prev_id = y   (ref :346-351)
// log_softmax is monotone per frame, so the argmax of log-probs needs no normalisation pass: the kernel
// reads every fp32 score exactly once (V*4 algorithmic bytes per frame) with 128-bit streaming loads.
//
// Grid: one CTA per (stream, 32-frame segment). Each warp reduces whole frames (coalesced row reads); the
// per-frame ids go to a small scratch; the LAST CTA of a stream to finish (atomic ticket) runs the
// order-dependent collapse for that stream with warp ballots, so no second launch is needed.
#include "k2b_internal.h"

namespace k2b {

namespace {

constexpr int kSeg = 32;       // frames per CTA
constexpr int kWarps = 8;

struct Arg {
  float v;
  int i;  // -1: no non-NaN value seen
};

__device__ __forceinline__ void take(Arg& b, float v, int i) {
  // .NET Max<float> orders NaN below every number; strict '>' keeps the first of equal values.
  if (v == v && (b.i < 0 || v > b.v)) { b.v = v; b.i = i; }
}

__device__ __forceinline__ Arg pick(const Arg& a, const Arg& b) {
  if (b.i < 0) return a;
  if (a.i < 0) return b;
  return (a.v > b.v || (a.v == b.v && a.i < b.i)) ? a : b;
}

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__global__ void __launch_bounds__(kWarps * 32)
ctc_greedy_kernel(const float* __restrict__ logp, int B, int T, int V, int blank, int nseg,
                  const int32_t* __restrict__ frame_offset, int64_t* __restrict__ prev_inout,
                  int64_t* __restrict__ tokens, int32_t* __restrict__ ts, int32_t* __restrict__ n_out,
                  int32_t* __restrict__ trailing_inout, int cap, int32_t* __restrict__ ybuf,
                  int32_t* __restrict__ ticket) {
  const int b = blockIdx.x / nseg, seg = blockIdx.x - b * nseg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_end = min(T, (seg + 1) * kSeg);

  for (int t = seg * kSeg + warp; t < t_end; t += kWarps) {
    const float* row = logp + ((size_t)b * T + t) * V;
    // rows are only 4-byte aligned in general (V = 5537): scalar head up to the next 16-byte boundary
    const int head = min(V, (int)(((16u - (unsigned)((uintptr_t)row & 15u)) & 15u) >> 2));
    const int nvec = (V - head) >> 2;
    const int tail0 = head + (nvec << 2);
    Arg best{0.f, -1};
    if (lane < head) take(best, __ldg(row + lane), lane);
    const float4* vp = reinterpret_cast<const float4*>(row + head);
    int q = lane;
    for (; q + 96 < nvec; q += 128) {   // 4 independent 128-bit loads in flight per lane
      const float4 x0 = ld_stream(vp + q), x1 = ld_stream(vp + q + 32), x2 = ld_stream(vp + q + 64),
                   x3 = ld_stream(vp + q + 96);
      int i = head + 4 * q;
      take(best, x0.x, i); take(best, x0.y, i + 1); take(best, x0.z, i + 2); take(best, x0.w, i + 3);
      i += 128;
      take(best, x1.x, i); take(best, x1.y, i + 1); take(best, x1.z, i + 2); take(best, x1.w, i + 3);
      i += 128;
      take(best, x2.x, i); take(best, x2.y, i + 1); take(best, x2.z, i + 2); take(best, x2.w, i + 3);
      i += 128;
      take(best, x3.x, i); take(best, x3.y, i + 1); take(best, x3.z, i + 2); take(best, x3.w, i + 3);
    }
    for (; q < nvec; q += 32) {
      const float4 x = ld_stream(vp + q);
      const int i = head + 4 * q;
      take(best, x.x, i); take(best, x.y, i + 1); take(best, x.z, i + 2); take(best, x.w, i + 3);
    }
    if (tail0 + lane < V) take(best, __ldg(row + tail0 + lane), tail0 + lane);
    // a lane's indices are not monotone across head/body/tail only in the sense head < body < tail: fine.
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      Arg o;
      o.v = __shfl_xor_sync(0xffffffffu, best.v, s);
      o.i = __shfl_xor_sync(0xffffffffu, best.i, s);
      best = pick(best, o);
    }
    if (lane == 0) ybuf[(size_t)b * T + t] = best.i < 0 ? 0 : best.i;  // all-NaN frame -> index 0
  }

  // ---- ticket: the last CTA of this stream collapses ------------------------------------------
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ticket[b], 1) == nseg - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last || warp != 0) return;
  __threadfence();

  const int off = frame_offset != nullptr ? frame_offset[b] : 0;
  int carry = prev_inout != nullptr ? (int)prev_inout[b] : -1;
  int base = 0, last_nonblank = -1;
  const int32_t* yb = ybuf + (size_t)b * T;
  // The collapse is a serial tail behind the last segment of every stream: its loads (one L2 round trip each) are issued eight
  // 32-frame groups at a time, so a 250-frame utterance pays ONE round trip instead of eight dependent ones.
  for (int g0 = 0; g0 < T; g0 += 256) {
    int ys[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = g0 + 32 * u + lane;
      ys[u] = t < T ? __ldcg(yb + t) : blank;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c0 = g0 + 32 * u;
      if (c0 >= T) break;                          // warp-uniform
      const int t = c0 + lane;
      const bool in = t < T;
      const int y = ys[u];
      int p = __shfl_up_sync(0xffffffffu, y, 1);
      if (lane == 0) p = carry;
      const bool emit = in && y != blank && y != p;
      const unsigned em = __ballot_sync(0xffffffffu, emit);
      const unsigned nb = __ballot_sync(0xffffffffu, in && y != blank);
      if (emit) {
        const int pos = base + __popc(em & ((1u << lane) - 1u));
        if (pos < cap) {
          tokens[(size_t)b * cap + pos] = y;
          ts[(size_t)b * cap + pos] = t + off;
        }
      }
      base += __popc(em);
      if (nb) last_nonblank = c0 + 31 - __clz(nb);
      carry = __shfl_sync(0xffffffffu, y, min(31, T - 1 - c0));
    }
  }
  if (lane == 0) {
    n_out[b] = base;
    if (trailing_inout != nullptr)
      trailing_inout[b] = last_nonblank < 0 ? trailing_inout[b] + T : T - 1 - last_nonblank;
    if (prev_inout != nullptr && T > 0) prev_inout[b] = carry;
    ticket[b] = 0;  // ready for the next launch
  }
}

}  // namespace

int32_t ctc_greedy_dev(k2b_handle* h, const float* logp, int B, int T, int V, int blank,
                       const int32_t* frame_offset, int64_t* prev_inout, int64_t* tokens, int32_t* ts,
                       int32_t* n_out, int32_t* trailing_inout, int cap) {
  if (B <= 0) return K2B_OK;
  if (T <= 0) {
    K2B_CUDA(h, cudaMemsetAsync(n_out, 0, sizeof(int32_t) * B, h->stream));
    return K2B_OK;
  }
  // scratch: tickets [B] int32 first (zero between launches: the collapsing CTA resets its own), then ybuf [B*T]
  const size_t ybuf_off = (sizeof(int32_t) * (size_t)B + 255) & ~size_t(255);
  const size_t need = ybuf_off + sizeof(int32_t) * (size_t)B * T;
  if (h->ws_ctc.bytes < need) {
    K2B_TRY(ensure(h, h->ws_ctc, need));
    K2B_CUDA(h, cudaMemsetAsync(h->ws_ctc.p, 0, h->ws_ctc.bytes, h->stream));
  }
  int32_t* ticket = static_cast<int32_t*>(h->ws_ctc.p);
  int32_t* ybuf = reinterpret_cast<int32_t*>(static_cast<char*>(h->ws_ctc.p) + ybuf_off);
  const int nseg = (T + kSeg - 1) / kSeg;
  const long long nblk = (long long)B * nseg;
  if (nblk > 0x7fffffffLL) return fail(h, K2B_ERR_INVALID, "ctc_greedy: B*ceil(T/32) exceeds the grid limit");
  prof_begin(h);
  ctc_greedy_kernel<<<(unsigned)nblk, kWarps * 32, 0, h->stream>>>(logp, B, T, V, blank, nseg, frame_offset, prev_inout,
                                                                  tokens, ts, n_out, trailing_inout, cap, ybuf, ticket);
  prof_end(h);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

}  // namespace k2b
