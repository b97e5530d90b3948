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
// Two forms (ctc_greedy_dev picks by input size):
//  * frames + collapse (inputs of 12 MB and more): `ctc_frames_kernel`, a PERSISTENT grid - two CTAs of eight warps per SM, frames
//    handed out eight at a time (one per warp) from a global counter whose next value is fetched one round ahead - in which a warp
//    reduces a whole frame with ALL of its 128-bit streaming loads in flight at once (16 per lane at V = 2000: one memory round trip
//    per frame) and posts nothing but the frame's id; `ctc_collapse_kernel` (one warp per stream: ballots for the order-dependent
//    blank / repeat collapse) is chained to it by a programmatic dependent launch and is resident, waiting, when the last frame is
//    reduced. cfg5's shard (128 streams = 256 MB): 46 us = 5.6 TB/s (torch.sum over the same bytes: 48 us); 1024 streams: 288 us =
//    7.1 TB/s (a pure read runs above the copy figure MEASURED_PEAKS.json holds).
//  * one kernel (small inputs, where a second launch costs more than it saves): `ctc_greedy_kernel`, the same hand-out with eight
//    loads per lane and batch, every frame posting a ticket for its stream with a release-atomic whose RESULT is looked at one frame
//    later; the warp that learns it posted a stream's last ticket runs the collapse. Round 2 measured what that costs at 256 MB
//    (62 us): the release is a MEMBAR.ALL.GPU on every frame's chain, the loop state spills at 64 registers, a frame is two
//    dependent round trips, and 4736 resident warps x 4 KB in flight make a frame - and therefore the tail - ~9 us long.
// The fold costs 2.5 instructions per score (three FMNMX per 16-byte group - they skip NaN like .NET's Max - one strict compare, and
// predicated copies of the winning group); a frame whose maximum is -inf or NaN-only is re-done by an exact slow path.
#include "k2b_internal.h"
#include "sm100_ptx.cuh"

namespace k2b {

namespace {

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

__device__ __forceinline__ Arg warp_pick(Arg best) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    Arg o;
    o.v = __shfl_xor_sync(0xffffffffu, best.v, s);
    o.i = __shfl_xor_sync(0xffffffffu, best.i, s);
    best = pick(best, o);
  }
  return best;
}

// exact for every input (NaN, -inf), about ten instructions per score: only frames the fast fold cannot decide come here
__device__ __noinline__ int frame_argmax_slow(const float* __restrict__ row, int V, int lane) {
  Arg best{0.f, -1};
  for (int i = lane; i < V; i += 32) take(best, __ldg(row + i), i);
  best = warp_pick(best);
  return best.i < 0 ? 0 : best.i;      // all-NaN frame -> index 0
}

// First index of the maximum of one frame (one warp). Strict '>' in index order keeps the first of equal maxima inside a lane,
// `pick` across lanes. A frame whose maximum is -inf (or that holds nothing but NaN) is not decided by the fold (-inf > -inf is
// false) and takes the exact slow path - log-probs of real models never do.
template <int NB>      // 128-bit loads in flight per lane and batch
__device__ __forceinline__ int frame_argmax(const float* __restrict__ row, int V, int lane) {
  // rows are only 4-byte aligned in general (V = 5537): scalar head up to the next 16-byte boundary
  const int head = min(V, (int)(((16u - (unsigned)((uintptr_t)row & 15u)) & 15u) >> 2));
  const int nvec = (V - head) >> 2;
  const int tail0 = head + (nvec << 2);
  Arg best{0.f, -1};
  if (lane < head) take(best, __ldg(row + lane), lane);
  const float4* vp = reinterpret_cast<const float4*>(row + head);
  float bv = best.i >= 0 ? best.v : -INFINITY;
  int bq = -1;
  float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
#define K2B_FOLD(v4_, qq) do { const float m_ = fmaxf(fmaxf((v4_).x, (v4_).y), fmaxf((v4_).z, (v4_).w)); const bool u_ = m_ > bv; \
    bv = u_ ? m_ : bv; bq = u_ ? (qq) : bq; bx.x = u_ ? (v4_).x : bx.x; bx.y = u_ ? (v4_).y : bx.y; bx.z = u_ ? (v4_).z : bx.z; \
    bx.w = u_ ? (v4_).w : bx.w; } while (0)
  int qb = 0;                          // warp-uniform batch base: every batch is one round trip for the whole warp
  for (; qb + 32 * NB <= nvec; qb += 32 * NB) {   // NB independent 128-bit loads in flight per lane
    float4 x[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) x[u] = ld_stream(vp + qb + lane + 32 * u);
#pragma unroll
    for (int u = 0; u < NB; ++u) K2B_FOLD(x[u], qb + lane + 32 * u);
  }
  if (qb < nvec) {                     // the rest of the row: up to NB more loads, issued together as well
    float4 x[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) if (qb + lane + 32 * u < nvec) x[u] = ld_stream(vp + qb + lane + 32 * u);
#pragma unroll
    for (int u = 0; u < NB; ++u) if (qb + lane + 32 * u < nvec) K2B_FOLD(x[u], qb + lane + 32 * u);
  }
#undef K2B_FOLD
  if (bq >= 0) {                       // the body beat the head: first position of bv inside its group
    best.v = bv;
    best.i = head + 4 * bq + (bx.x == bv ? 0 : (bx.y == bv ? 1 : (bx.z == bv ? 2 : 3)));
  }
  if (tail0 + lane < V) take(best, __ldg(row + tail0 + lane), tail0 + lane);
  best = warp_pick(best);
  if (best.i < 0 || best.v == -INFINITY) return frame_argmax_slow(row, V, lane);   // warp-uniform
  return best.i;
}

// the order-dependent part, one warp per stream: collapse repeats, drop blanks, count trailing blanks
__device__ __noinline__ void collapse_stream(int b, int T, int blank, int lane, const int32_t* __restrict__ frame_offset,
                                             int64_t* __restrict__ prev_inout, int64_t* __restrict__ tokens, int32_t* __restrict__ ts,
                                             int32_t* __restrict__ n_out, int32_t* __restrict__ trailing_inout, int cap,
                                             const int32_t* __restrict__ ybuf) {
  const int off = frame_offset != nullptr ? frame_offset[b] : 0;
  int carry = prev_inout != nullptr ? (int)prev_inout[b] : -1;
  int base = 0, last_nonblank = -1;
  const int32_t* yb = ybuf + (size_t)b * T;
  // the loads (one L2 round trip each) are issued eight 32-frame groups at a time: a 250-frame utterance pays ONE round trip
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
  }
}

__device__ __forceinline__ int atom_add_release(int32_t* p, int v) {
  int old;
  asm volatile("atom.release.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

__global__ void __launch_bounds__(kWarps * 32, 4)
ctc_greedy_kernel(const float* __restrict__ logp, int B, int T, int V, int blank,
                  const int32_t* __restrict__ frame_offset, int64_t* __restrict__ prev_inout,
                  int64_t* __restrict__ tokens, int32_t* __restrict__ ts, int32_t* __restrict__ n_out,
                  int32_t* __restrict__ trailing_inout, int cap, int32_t* __restrict__ ybuf,
                  int32_t* __restrict__ ticket, int32_t* __restrict__ sched, int G) {
  __shared__ int s_base[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)B * T;
  // G frames per warp and grab: 1 when the input is only a few frames per warp (the tail must stay short), up to 4 when it is
  // large (the warps of a CTA then meet at the barrier every fourth frame only)
  if (threadIdx.x == 0) s_base[0] = atomicAdd(&sched[0], kWarps * G);
  int pend_b = -1, pend_old = 0;        // lane 0: the ticket posted for the previous frame (its result is looked at a frame later)
  auto retire = [&]() {
    int fin = (lane == 0 && pend_b >= 0 && pend_old == T - 1) ? pend_b : -1;
    fin = __shfl_sync(0xffffffffu, fin, 0);
    if (fin >= 0) {                      // this warp posted the last ticket of stream `fin`: every frame of it has been reduced
      __threadfence();
      collapse_stream(fin, T, blank, lane, frame_offset, prev_inout, tokens, ts, n_out, trailing_inout, cap, ybuf);
      if (lane == 0) ticket[fin] = 0;    // ready for the next launch
    }
    pend_b = -1;
  };
  for (int round = 0;; ++round) {
    __syncthreads();
    const long long base = s_base[round & 1];
    if (base >= total) break;                                                 // CTA-uniform
    int next = 0;
    if (threadIdx.x == 0) next = atomicAdd(&sched[0], kWarps * G);            // next round's frames: fetched behind this round's loads
    for (int g = 0; g < G; ++g) {
      const long long f = base + warp + (long long)g * kWarps;
      if (f >= total) break;
      const int b = (int)(f / T);
      const int y = frame_argmax<8>(logp + (size_t)f * V, V, lane);
      retire();
      if (lane == 0) {
        ybuf[f] = y;
        pend_old = atom_add_release(&ticket[b], 1);
        pend_b = b;
      }
    }
    if (threadIdx.x == 0) s_base[(round + 1) & 1] = next;
  }
  retire();
  // the last CTA to leave re-arms the frame counter (every CTA has made its final, failing, grab by then)
  if (threadIdx.x == 0 && atomicAdd(&sched[1], 1) == (int)gridDim.x - 1) {
    sched[0] = 0;
    sched[1] = 0;
  }
}

// ---- round 2, second session: the per-frame work and the order-dependent part as TWO kernels chained by a programmatic dependent
// launch. What ncu showed on the one-kernel version at cfg5's shard (128 streams = 256 MB, 62 us): the release-atomic ticket of
// every frame is a MEMBAR.ALL.GPU on the warp's chain (stall_membar 2.6 warps per issue cycle), the loop state spills (15 STL +
// 18 LDL per frame at 64 registers), a frame is two dependent load round trips, and with 4736 resident warps x 4 KB in flight
// (19 MB) the queueing delay makes a frame ~9 us long - which is also the length of the tail. Here a frame is ONE round trip (all
// of its 128-bit loads - 16 per lane at V = 2000 - are issued before the first fold; two CTAs per SM, 128 registers, no spills),
// the frame kernel posts nothing but the per-frame id (no ticket, no fence, no 64-bit division), and the collapse kernel (one warp
// per stream) is launched with programmatic stream serialisation: it is resident and waiting when the last frame is reduced.
__global__ void __launch_bounds__(kWarps * 32, 2)
ctc_frames_kernel(const float* __restrict__ logp, long long total, int V, int32_t* __restrict__ ybuf, int32_t* __restrict__ sched) {
  __shared__ int s_base[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::griddep_launch_dependents();     // the collapse grid may take its SM slots as soon as they free up; it waits for this grid
  if (threadIdx.x == 0) s_base[0] = atomicAdd(&sched[0], kWarps);
  for (int round = 0;; ++round) {
    __syncthreads();
    const long long base = s_base[round & 1];
    if (base >= total) break;                                                 // CTA-uniform
    int next = 0;
    if (threadIdx.x == 0) next = atomicAdd(&sched[0], kWarps);                // next round's frames: fetched behind this round's loads
    const long long f = base + warp;
    if (f < total) {
      const int y = frame_argmax<16>(logp + (size_t)f * V, V, lane);
      if (lane == 0) ybuf[f] = y;
    }
    if (threadIdx.x == 0) s_base[(round + 1) & 1] = next;
  }
  // the last CTA to leave re-arms the frame counter (every CTA has made its final, failing, grab by then)
  if (threadIdx.x == 0 && atomicAdd(&sched[1], 1) == (int)gridDim.x - 1) {
    sched[0] = 0;
    sched[1] = 0;
  }
}

__global__ void __launch_bounds__(kWarps * 32)
ctc_collapse_kernel(int B, int T, int blank, const int32_t* __restrict__ frame_offset, int64_t* __restrict__ prev_inout,
                    int64_t* __restrict__ tokens, int32_t* __restrict__ ts, int32_t* __restrict__ n_out,
                    int32_t* __restrict__ trailing_inout, int cap, const int32_t* __restrict__ ybuf) {
  const int b = blockIdx.x * kWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  ptx::griddep_wait();                  // every frame id of the frame kernel is written and visible
  if (b < B) collapse_stream(b, T, blank, lane, frame_offset, prev_inout, tokens, ts, n_out, trailing_inout, cap, ybuf);
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
  // scratch (zero between launches: the kernel re-arms it): frame counter + exit counter, tickets [B] int32, then ybuf [B*T]
  const size_t tick_off = 256;
  const size_t ybuf_off = tick_off + ((sizeof(int32_t) * (size_t)B + 255) & ~size_t(255));
  const size_t need = ybuf_off + sizeof(int32_t) * (size_t)B * T;
  if ((long long)B * T + kWarps * 4 * 8192 > 0x7fffffffLL) return fail(h, K2B_ERR_INVALID, "ctc_greedy: B*T exceeds the frame counter");
  if (h->ws_ctc.bytes < need) {
    K2B_TRY(ensure(h, h->ws_ctc, need));
    K2B_CUDA(h, cudaMemsetAsync(h->ws_ctc.p, 0, h->ws_ctc.bytes, h->stream));
  }
  int32_t* sched = static_cast<int32_t*>(h->ws_ctc.p);
  int32_t* ticket = reinterpret_cast<int32_t*>(static_cast<char*>(h->ws_ctc.p) + tick_off);
  int32_t* ybuf = reinterpret_cast<int32_t*>(static_cast<char*>(h->ws_ctc.p) + ybuf_off);
  const long long rounds = ((long long)B * T + kWarps - 1) / kWarps;
  const double bytes = (double)B * T * V * sizeof(float);
  // small inputs (measured crossover between 8 and 16 MB, tools/run_ctc_shard.py): one kernel with tickets - a second launch costs
  // more than the fences; everything else: frames + collapse as two kernels chained by a programmatic dependent launch
  const bool one_kernel = h->opt_ctc_one_kernel >= 0 ? h->opt_ctc_one_kernel == 1 : bytes < 12e6;
  if (!one_kernel) {
    static int per_sm2 = 0;
    if (per_sm2 == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, ctc_frames_kernel, kWarps * 32, 0) != cudaSuccess || per_sm2 < 1)) {
      cudaGetLastError();
      per_sm2 = 2;
    }
    const long long nblk2 = rounds < (long long)per_sm2 * h->sm_count ? rounds : (long long)per_sm2 * h->sm_count;
    prof_begin(h);
    ctc_frames_kernel<<<(unsigned)nblk2, kWarps * 32, 0, h->stream>>>(logp, (long long)B * T, V, ybuf, sched);
    K2B_LAUNCH_CHECK(h);
    K2B_CUDA(h, launch_pdl(ctc_collapse_kernel, dim3((unsigned)((B + kWarps - 1) / kWarps)), dim3(kWarps * 32), 0, h->stream, B, T, blank,
                           frame_offset, prev_inout, tokens, ts, n_out, trailing_inout, cap, (const int32_t*)ybuf));
    prof_end(h);
    h->launches++;
    return K2B_OK;
  }
  static int per_sm = 0;
  if (per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ctc_greedy_kernel, kWarps * 32, 0) != cudaSuccess || per_sm < 1)) {
    cudaGetLastError();
    per_sm = 4;
  }
  const long long nblk = rounds < (long long)per_sm * h->sm_count ? rounds : (long long)per_sm * h->sm_count;
  const long long per_warp = (long long)B * T / (nblk * kWarps);
  const int G = per_warp >= 32 ? 4 : (per_warp >= 16 ? 2 : 1);
  prof_begin(h);
  ctc_greedy_kernel<<<(unsigned)nblk, kWarps * 32, 0, h->stream>>>(logp, B, T, V, blank, frame_offset, prev_inout,
                                                                  tokens, ts, n_out, trailing_inout, cap, ybuf, ticket, sched, G);
  prof_end(h);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

}  // namespace k2b
