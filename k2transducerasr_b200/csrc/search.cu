// Transducer search loops on the device: greedy_search (offline single / batch / online chunk) and
// modified_beam_search. The time loop is a sequence of launches on the handle's stream, three per frame:
//   1. decoder GEMM   x = tanh(enc[s,t] + dec_proj(relu(tab0[y0] + tab1[y1])))          (gemm_*.cu)
//   2. joiner GEMM    logits tile = x * out_w^T + out_b  -> per-(row, vocab tile) partials only
//   3. select         greedy: fold argmax partials, append, shift context
//                     beam:   log_softmax constants, per-stream top-K over K*V, hypothesis merge
// Nothing crosses PCIe inside the loop and the logits are never written to HBM.
#include <math.h>
#include <stdlib.h>

#include "k2b_internal.h"
#include "sm100_ptx.cuh"
#include "beam_merge.cuh"

namespace k2b {

namespace {


// ------------------------------------------------------------------------------------------------
// greedy
// ------------------------------------------------------------------------------------------------
__global__ void greedy_init_kernel(int B, int blank, const int64_t* __restrict__ hyp_in, int32_t* __restrict__ ctx,
                                   int32_t* __restrict__ n_out, int32_t* __restrict__ flag) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *flag = 0;
  if (b >= B) return;
  if (hyp_in != nullptr) {           // online: context = OnlineStream.Hyp (ref OnlineRecognizer.cs:109,125)
    ctx[2 * b] = (int32_t)hyp_in[2 * b];
    ctx[2 * b + 1] = (int32_t)hyp_in[2 * b + 1];
  } else {                           // offline: {-1, blank} (ref OfflineRecognizer.cs:105, :202)
    ctx[2 * b] = -1;
    ctx[2 * b + 1] = blank;
  }
  n_out[b] = 0;
}

// One thread per stream: fold the per-tile argmax partials in index order (ties / NaN -> larger index,
// ref OfflineRecognizer.cs:153), then the emission test (ref :161 offline, OnlineRecognizer.cs:181 online).
__global__ void greedy_select_kernel(int B, int nt, const float* __restrict__ pval, const int32_t* __restrict__ pidx,
                                     const int32_t* __restrict__ pnan, int32_t* __restrict__ ctx,
                                     int64_t* __restrict__ tokens, int32_t* __restrict__ ts, int32_t* __restrict__ n_out,
                                     int cap, int t, int blank, int unk, int extra_mask, int max_sym,
                                     int32_t* __restrict__ flag, const int32_t* __restrict__ lens, int sub, int32_t* __restrict__ spf) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (lens != nullptr && t >= lens[b]) return;      // ragged batch: this stream has ended
  // max_sym_per_frame > 1 (ref OfflineRecognizer.cs:127-179): evaluation `sub` of frame t only concerns the streams that emitted
  // on each of the `sub` evaluations before it
  if (spf != nullptr) {
    if (sub == 0) spf[b] = 0;
    else if (spf[b] != sub) return;
  }
  float bv = 0.f;
  int bi = -1, bn = 0;
  for (int i = 0; i < nt; ++i) {
    const size_t o = (size_t)b * nt + i;
    const float v = pval[o];
    const int idx = pidx[o], nn = pnan[o];
    if (idx < 0) continue;
    if (bi < 0 || nn) { bv = v; bi = idx; bn = nn | bn; continue; }
    if (!(bv > v)) { bv = v; bi = idx; }
  }
  (void)bn;
  const int n = n_out[b];
  if (n >= max_sym) return;          // single-stream loop stops at 1000 symbols (ref OfflineRecognizer.cs:122,127)
  const int y = bi;
  if (y != blank && y != unk && y != extra_mask) {
    if (n < cap) {
      tokens[(size_t)b * cap + n] = y;
      ts[(size_t)b * cap + n] = t;
    }
    n_out[b] = n + 1;
    ctx[2 * b] = ctx[2 * b + 1];
    ctx[2 * b + 1] = y;
    *flag = 1;                       // Q6: somebody in the batch emitted
    if (spf != nullptr) spf[b] = sub + 1;
  }
}

__global__ void greedy_finish_online_kernel(int B, const int32_t* __restrict__ ctx, int64_t* __restrict__ hyp_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  hyp_out[2 * b] = ctx[2 * b];       // ref OnlineRecognizer.cs:208
  hyp_out[2 * b + 1] = ctx[2 * b + 1];
}

// ------------------------------------------------------------------------------------------------
// modified_beam_search
// ------------------------------------------------------------------------------------------------

__global__ void beam_init_kernel(int B, int K, int blank, BeamState s0, BeamState s1) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B * K) return;
  const int slot = n % K;
  s0.ctx[2 * n] = -1; s0.ctx[2 * n + 1] = blank;
  s1.ctx[2 * n] = -1; s1.ctx[2 * n + 1] = blank;
  s0.lp[n] = slot == 0 ? 0.f : -INFINITY;
  s0.len[n] = 2;
  s0.hash[n] = kHashSeed;
  if (s0.cst != nullptr) { s0.cst[n] = 0; s1.cst[n] = 0; }
  if (slot == 0) s0.nlive[n / K] = 1;
}

__device__ __forceinline__ bool better(float v, int i, float ev, int ei) {
  return v > ev || (v == ev && i > ei);
}

// One warp per stream ("hyp_merge"): log_softmax constants per live hypothesis from the tile partials,
// per-stream top-K over the K*V extensions (value desc, flat index desc), extension, dedupe by
// token-sequence hash with log-add, compaction in insertion order, back-pointer record.
// Branch-free throughout (predicated selects, REDUX max for the arg-best rounds): the lanes of a warp never diverge.

__global__ void __launch_bounds__(128)
beam_select_kernel(int B, int K, int V, int nt, int T, int t, int blank, int unk, int mask3,
                   const float* __restrict__ part_m, const float* __restrict__ part_s,
                   const float* __restrict__ part_tv, const int32_t* __restrict__ part_ti,
                   BeamState in, BeamState out, int32_t* __restrict__ bp, const int32_t* __restrict__ lens,
                   const int32_t* __restrict__ cg_next, const float* __restrict__ cg_delta) {
  __shared__ float s_mx[4][kMaxBeam], s_ls[4][kMaxBeam], s_lp[4][kMaxBeam];
  constexpr int kNone = (int)0x80000000;
  const unsigned full = 0xffffffffu;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * 4 + wib;
  k2b::ptx::griddep_launch_dependents();
  k2b::ptx::griddep_wait();                             // everything read below is the output of the kernels before this one
  if (s >= B) return;
  const int nl = in.nlive[s];
  if (lens != nullptr && t >= lens[s]) {            // ragged batch: past the end of this stream, the hypotheses are frozen
    if (lane < K) {
      const size_t o = (size_t)s * K + lane;
      out.ctx[2 * o] = in.ctx[2 * o]; out.ctx[2 * o + 1] = in.ctx[2 * o + 1];
      out.lp[o] = in.lp[o]; out.len[o] = in.len[o]; out.hash[o] = in.hash[o];
      if (cg_next != nullptr) out.cst[o] = in.cst[o];
      bp[((size_t)s * T + t) * K + lane] = lane < nl ? (lane << 28) : 0;
    }
    if (lane == 0) out.nlive[s] = nl;
    return;
  }

  // log_softmax constants. Every (hypothesis, tile) pair is one work item, lane-strided, with all loads of a batch of four
  // items in flight at once (the partials sit in L2 / HBM: latency, not bandwidth, is what this kernel pays for):
  // pass 1 folds the tile maxima per hypothesis (REDUX max), pass 2 the rescaled sums (butterfly).
  const int npair = nl * nt;
  const size_t prow0 = (size_t)s * K * nt;                 // pair p of this stream = element prow0 + p of part_m / part_s
  int mk[kMaxBeam];
#pragma unroll
  for (int h = 0; h < kMaxBeam; ++h) mk[h] = kNone;
  for (int p0 = 0; p0 < npair; p0 += 128) {
    float pm[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int p = p0 + 32 * u + lane; pm[u] = p < npair ? part_m[prow0 + p] : -INFINITY; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + 32 * u + lane;
      const int hp = p / nt, key = p < npair ? fkey_s(pm[u]) : kNone;
#pragma unroll
      for (int h = 0; h < kMaxBeam; ++h) mk[h] = (hp == h) ? max(mk[h], key) : mk[h];
    }
  }
  float mxh[kMaxBeam], sumh[kMaxBeam];
#pragma unroll
  for (int h = 0; h < kMaxBeam; ++h) {
    mxh[h] = (h < K) ? funkey_s(__reduce_max_sync(full, mk[h])) : -INFINITY;
    sumh[h] = 0.f;
  }
  for (int p0 = 0; p0 < npair; p0 += 128) {
    float pm[4], ps[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + 32 * u + lane;
      pm[u] = p < npair ? part_m[prow0 + p] : -INFINITY;
      ps[u] = p < npair ? part_s[prow0 + p] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int hp = (p0 + 32 * u + lane) / nt;
      float mh = -INFINITY;
#pragma unroll
      for (int h = 0; h < kMaxBeam; ++h) mh = (hp == h) ? mxh[h] : mh;
      const float term = (pm[u] > -INFINITY) ? ps[u] * __expf(pm[u] - mh) : 0.f;
#pragma unroll
      for (int h = 0; h < kMaxBeam; ++h) sumh[h] += (hp == h) ? term : 0.f;
    }
  }
#pragma unroll
  for (int h = 0; h < kMaxBeam; ++h) {
    if (h < K) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) sumh[h] += __shfl_xor_sync(full, sumh[h], o);
      if (lane == 0 && h < nl) {
        s_mx[wib][h] = mxh[h];
        s_ls[wib][h] = __logf(sumh[h]);
        s_lp[wib][h] = in.lp[(size_t)s * K + h];
      }
    }
  }
  __syncwarp();

  // lane-local top-K (sorted, best first) of this lane's share of the candidates: predicated insertion network,
  // candidates loaded four at a time
  int tk[kMaxBeam], tf[kMaxBeam];
#pragma unroll
  for (int i = 0; i < kMaxBeam; ++i) { tk[i] = kNone; tf[i] = -1; }
  const int per_h = nt * K;
  const int total = nl * per_h;
  const size_t base = (size_t)s * K * per_h;
  for (int c0 = 0; c0 < total; c0 += 128) {
    int cidx[4];
    float cval[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 32 * u + lane;
      cidx[u] = c < total ? part_ti[base + c] : -1;
      cval[u] = c < total ? part_tv[base + c] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 32 * u + lane;
      const int h = min(c / per_h, kMaxBeam - 1);
      // same operation order as log_softmax(x) + lp : ((x - max) - log(sum)) + lp
      const float v = ((cval[u] - s_mx[wib][h]) - s_ls[wib][h]) + s_lp[wib][h];
      const bool okc = (cidx[u] >= 0) & (v == v);
      int key = okc ? fkey_s(v) : kNone;
      int f = okc ? h * V + cidx[u] : -1;
#pragma unroll
      for (int i = 0; i < kMaxBeam; ++i) {
        if (i < K) {
          const bool b = (key > tk[i]) | ((key == tk[i]) & (f > tf[i]));
          const int nk = b ? tk[i] : key, nf = b ? tf[i] : f;
          tk[i] = b ? key : tk[i]; tf[i] = b ? f : tf[i];
          key = nk; f = nf;
        }
      }
    }
  }

  // K rounds of warp arg-best over the heads (REDUX on the key, then on the flat index among the ties); lane r keeps winner r
  float my_v = -INFINITY;
  int my_f = -1;
#pragma unroll
  for (int r = 0; r < kMaxBeam; ++r) {
    if (r < K) {
      const int wk = __reduce_max_sync(full, tk[0]);
      const int wf = __reduce_max_sync(full, (tk[0] == wk) ? tf[0] : -1);
      const bool pop = (tf[0] == wf) & (wf >= 0);      // flat indices are unique: exactly one lane pops
#pragma unroll
      for (int i = 0; i + 1 < kMaxBeam; ++i) { tk[i] = pop ? tk[i + 1] : tk[i]; tf[i] = pop ? tf[i + 1] : tf[i]; }
      tk[kMaxBeam - 1] = pop ? kNone : tk[kMaxBeam - 1]; tf[kMaxBeam - 1] = pop ? -1 : tf[kMaxBeam - 1];
      if (lane == r) { my_v = wf >= 0 ? funkey_s(wk) : -INFINITY; my_f = wf; }
    }
  }

  // lane r < K: the r-th extension in rank order
  const bool cand = lane < K && my_f >= 0;
  int par = 0, tok = -1, c0 = -1, c1 = blank, ln = 2, cs = 0;
  uint64_t hs = kHashSeed;
  if (cand) {
    par = my_f / V;
    const int y = my_f - par * V;
    const size_t prow = (size_t)s * K + par;
    hs = in.hash[prow];
    ln = in.len[prow];
    c0 = in.ctx[2 * prow];
    c1 = in.ctx[2 * prow + 1];
    if (cg_next != nullptr) cs = in.cst[prow];
    if (y != blank && y != unk && y != mask3) {      // ys unchanged for blank / unk (/ the literal 1 of the online loop)
      tok = y;
      hs = hash_push(hs, y);
      ln += 1;
      c0 = c1;
      c1 = y;
      if (cg_next != nullptr) {                      // hot words: the boost is applied AFTER the top-K selection, to the extended hypothesis
        my_v += cg_delta[(size_t)cs * V + y];
        cs = cg_next[(size_t)cs * V + y];
      }
    }
  }
  // dedupe: first earlier lane holding the same token sequence
  int root = lane;
  for (int q = 0; q < K; ++q) {
    const uint64_t qh = __shfl_sync(0xffffffffu, hs, q);
    const int ql = __shfl_sync(0xffffffffu, ln, q);
    const int q0 = __shfl_sync(0xffffffffu, c0, q);
    const int q1 = __shfl_sync(0xffffffffu, c1, q);
    const int qc = __shfl_sync(0xffffffffu, (int)cand, q);
    if (cand && qc && q < lane && root == lane && qh == hs && ql == ln && q0 == c0 && q1 == c1) root = q;
  }
  // log-add the merged scores into their root, in insertion (rank) order
  float lp = my_v;
  for (int q = 0; q < K; ++q) {
    const int qroot = __shfl_sync(0xffffffffu, root, q);
    const float qv = __shfl_sync(0xffffffffu, my_v, q);
    const int qc = __shfl_sync(0xffffffffu, (int)cand, q);
    if (cand && qc && q != lane && qroot == lane) lp = logaddexp_f(lp, qv);
  }
  const bool is_root = cand && root == lane;
  const unsigned roots = __ballot_sync(0xffffffffu, is_root);
  const int nnew = __popc(roots);
  if (is_root) {
    const int slot = __popc(roots & ((1u << lane) - 1u));
    const size_t o = (size_t)s * K + slot;
    out.ctx[2 * o] = c0;
    out.ctx[2 * o + 1] = c1;
    out.lp[o] = lp;
    out.len[o] = ln;
    out.hash[o] = hs;
    if (cg_next != nullptr) out.cst[o] = cs;
    bp[((size_t)s * T + t) * K + slot] = (par << 28) | (tok + 1);
  }
  if (lane >= nnew && lane < K) {      // dead slots keep a valid context for the next decoder GEMM
    const size_t o = (size_t)s * K + lane;
    out.ctx[2 * o] = -1;
    out.ctx[2 * o + 1] = blank;
    out.lp[o] = -INFINITY;
    out.len[o] = 2;
    out.hash[o] = kHashSeed;
    if (cg_next != nullptr) out.cst[o] = 0;
    bp[((size_t)s * T + t) * K + lane] = 0;
  }
  if (lane == 0) out.nlive[s] = nnew;
}

// Fused frame step of the memoised-decoder path, one CTA per stream ("hyp_merge" + the next frame's joiner operand):
//   A. warp h <-> live hypothesis h: log_softmax constants from its tile partials, its K best extensions (all loads of a
//      hypothesis are issued before the first use: the partials sit in L2, the kernel pays latency, not bandwidth);
//   B. warp 0: the stream's top K over the K*K survivors (value desc, flat index desc), extension, dedupe by token-sequence hash
//      with log-add in rank order, compaction, back-pointer record - the code of beam_select_kernel;
//   C. all threads: x[m,:] = tanh(enc[s,t+1] + decoder(ctx[m])) of the K new hypotheses from the memoised decoder table, written
//      as the bf16 hi / lo tile images the joiner's loader warp fetches (what joinin_table_kernel does in a launch of its own).
// Same results as beam_select_kernel + joinin_table_kernel; one launch and one round trip of the contexts through HBM less.
// Stand-alone launch of the frame step, one CTA per stream (beam_merge.cuh): same results as beam_select_kernel +
// joinin_table_kernel; one launch and one round trip of the contexts through HBM less.
template <int KB>
__global__ void __launch_bounds__(128)
beam_step_kernel(int B, int K, int V, int nt, int T, int t, int blank, int unk, int mask3,
                 const float* __restrict__ part_rec,
                 BeamState in, BeamState out, int32_t* __restrict__ bp, const int32_t* __restrict__ lens,
                 const float* __restrict__ dec_tab, const float* __restrict__ enc_next, long long enc_stride, int J,
                 uint8_t* __restrict__ x_img, long long* __restrict__ tl) {
  __shared__ float c_v[KB * KB];
  __shared__ int c_f[KB * KB];
  __shared__ int s_ctx[2 * KB];
  const int tid = threadIdx.x, s = blockIdx.x;
  k2b::ptx::griddep_launch_dependents();
  float4 pe0, pe1;            // the next frame of this stream does not depend on the kernels before this one: fetched ahead of the wait
  beam_merge_prefetch(tid, s, K, J, enc_next, enc_stride, &pe0, &pe1);
  if (tl != nullptr) {                                      // diagnostic timeline: [sm][8], slots 4..7 = first start, first wait-done,
    uint32_t smid;                                          // last merge-done, last end of the CTAs of this launch on that SM
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tl += (size_t)smid * 8;
    if (tid == 0) atomicMin(reinterpret_cast<long long*>(tl + 4), clock64());
  }
  k2b::ptx::griddep_wait();
  if (tl != nullptr && tid == 0) atomicMin(reinterpret_cast<long long*>(tl + 5), clock64());
  beam_merge_stream<KB>(tid, 1, s, K, V, nt, T, t, blank, unk, mask3, part_rec, in, out, bp, lens, dec_tab, enc_next, enc_stride, J, x_img,
                        pe0, pe1, c_v, c_f, s_ctx, tl);
}

// One warp per stream: pick argmax lp/len (first maximum in slot order), walk the back-pointers in
// shared memory, write tokens / timestamps in forward order.
__global__ void __launch_bounds__(32)
beam_backtrace_kernel(int B, int K, int T, BeamState fin, const int32_t* __restrict__ bp, int64_t* __restrict__ tokens,
                      int32_t* __restrict__ ts, int32_t* __restrict__ n_out, float* __restrict__ score, int cap,
                      const float* __restrict__ cg_resid) {
  extern __shared__ int32_t sm[];
  int32_t* trel = sm;                 // [T*K]
  int32_t* rtok = sm + (size_t)T * K; // [T]
  int32_t* rts = rtok + T;            // [T]
  const int s = blockIdx.x, lane = threadIdx.x;
  const int32_t* src = bp + (size_t)s * T * K;
  for (int i = lane; i < T * K; i += 32) trel[i] = src[i];
  __syncwarp();
  int n = 0;
  if (lane == 0) {
    const int nl = fin.nlive[s];
    int best = 0;
    float bn = -INFINITY;
    for (int q = 0; q < nl; ++q) {
      // hot words: the boost of a match that did not complete is revoked before the hypotheses are compared
      const float lpq = fin.lp[(size_t)s * K + q] - (cg_resid != nullptr ? cg_resid[fin.cst[(size_t)s * K + q]] : 0.f);
      const float norm = __fdiv_rn(lpq, (float)fin.len[(size_t)s * K + q]);
      if (q == 0 || norm > bn) { bn = norm; best = q; }
    }
    score[s] = fin.lp[(size_t)s * K + best] - (cg_resid != nullptr ? cg_resid[fin.cst[(size_t)s * K + best]] : 0.f);
    int slot = best;
    for (int t = T - 1; t >= 0; --t) {
      const int e = trel[t * K + slot];
      const int tok = (e & 0x0fffffff) - 1;
      if (tok >= 0) { rtok[n] = tok; rts[n] = t; ++n; }
      slot = (e >> 28) & 0xf;
    }
    n_out[s] = n;
  }
  n = __shfl_sync(0xffffffffu, n, 0);
  __syncwarp();
  for (int i = lane; i < n && i < cap; i += 32) {
    tokens[(size_t)s * cap + i] = rtok[n - 1 - i];
    ts[(size_t)s * cap + i] = rts[n - 1 - i];
  }
}

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

BeamState carve_state(char*& p, int B, int K) {
  const size_t N = (size_t)B * K;
  BeamState s;
  s.ctx = reinterpret_cast<int32_t*>(p);  p += align256(N * 2 * sizeof(int32_t));
  s.lp = reinterpret_cast<float*>(p);     p += align256(N * sizeof(float));
  s.len = reinterpret_cast<int32_t*>(p);  p += align256(N * sizeof(int32_t));
  s.hash = reinterpret_cast<uint64_t*>(p); p += align256(N * sizeof(uint64_t));
  s.nlive = reinterpret_cast<int32_t*>(p); p += align256((size_t)B * sizeof(int32_t));
  s.cst = reinterpret_cast<int32_t*>(p);  p += align256(N * sizeof(int32_t));
  return s;
}

size_t state_bytes(int B, int K) {
  const size_t N = (size_t)B * K;
  return align256(N * 8) + align256(N * 4) + align256(N * 4) + align256(N * 8) + align256((size_t)B * 4) + align256(N * 4);
}

}  // namespace

size_t beam_state_bytes(int B, int K) { return state_bytes(B, K); }
BeamStateView beam_state_view(k2b_handle* h, int B, int K, int which) {
  char* p = static_cast<char*>(h->ws_state.p) + (size_t)which * state_bytes(B, K);
  const BeamState s = carve_state(p, B, K);
  return BeamStateView{s.ctx, s.lp, s.len, reinterpret_cast<unsigned long long*>(s.hash), s.nlive, s.cst};
}

void prof_begin(k2b_handle* h) {
  if (!h->profile_on) return;
  ProfEvents& p = h->prof;
  if (p.used == p.start.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { h->profile_on = false; return; }
    p.start.push_back(a);
    p.stop.push_back(b);
  }
  cudaEventRecord(p.start[p.used], h->stream);
}

void prof_end(k2b_handle* h) {
  if (!h->profile_on) return;
  ProfEvents& p = h->prof;
  cudaEventRecord(p.stop[p.used], h->stream);
  p.used++;
}

int32_t greedy_dev(k2b_handle* h, const float* enc, int B, int T, int mode, bool online, int64_t* hyp_inout,
                   int64_t* tokens, int32_t* ts, int32_t* n_out, int cap) {
  const k2b_config& c = h->cfg;
  const int J = c.joiner_dim, V = c.vocab_size, D = c.decoder_dim;
  const bool tc = c.precision != K2B_PREC_FP32 && joiner_tc_supported(h);   // per-frame tcgen05 joiner (256-column tiles)
  const int nt = tc ? joiner_tc_tiles(h) : num_vocab_tiles(V);
  K2B_TRY(ensure(h, h->ws_x, sizeof(float) * (size_t)B * J));
  K2B_TRY(ensure(h, h->ws_part, (size_t)B * nt * 12));
  h->ll_clean_ptr = nullptr;             // (the persistent greedy kernel's tagged records share this buffer)
  K2B_TRY(ensure(h, h->ws_state, align256((size_t)B * 8) + 256 + align256((size_t)B * 4)));
  int32_t* ctx = static_cast<int32_t*>(h->ws_state.p);
  int32_t* flag = reinterpret_cast<int32_t*>(static_cast<char*>(h->ws_state.p) + align256((size_t)B * 8));
  // symbols per frame: only the single-stream loop of the reference has the notion (ref OfflineRecognizer.cs:19, :127-134); the
  // batch and online loops evaluate every frame once by construction
  const int nsub = (!online && mode != K2B_GREEDY_BATCH_COMPAT) ? h->max_sym_per_frame : 1;
  int32_t* spf = nsub > 1 ? reinterpret_cast<int32_t*>(static_cast<char*>(h->ws_state.p) + align256((size_t)B * 8) + 256) : nullptr;
  float* pval = static_cast<float*>(h->ws_part.p);
  int32_t* pidx = reinterpret_cast<int32_t*>(pval + (size_t)B * nt);
  int32_t* pnan = pidx + (size_t)B * nt;
  float* x = static_cast<float*>(h->ws_x.p);

  const int tb = 128, gb = (B + tb - 1) / tb;
  greedy_init_kernel<<<gb, tb, 0, h->stream>>>(B, c.blank_id, online ? hyp_inout : nullptr, ctx, n_out, flag);
  K2B_LAUNCH_CHECK(h);

  const int max_sym = (mode == K2B_GREEDY_SINGLE && !online) ? 1000 : 0x7fffffff;
  const int extra_mask = online ? 1 : -1;   // the literal `y != 1` of ref OnlineRecognizer.cs:181
  for (int t = 0; t < T; ++t)
  for (int sub = 0; sub < nsub; ++sub) {
    GemmArgs d;
    d.M = B; d.N = J; d.K = D;
    d.W = h->dec_w; d.bias = h->dec_b;
    d.ctx = ctx; d.tab0 = h->tab0; d.tab1 = h->tab1; d.V = V; d.neg_wrap = c.neg_id_mode == K2B_NEGID_WRAP;
    d.blank = c.blank_id;
    d.compat_flag = (mode == K2B_GREEDY_BATCH_COMPAT && !online) ? flag : nullptr;
    d.enc = enc + (size_t)t * J; d.enc_stride = (long long)T * J; d.rows_per_stream = 1;
    d.C = x;
    K2B_TRY(launch_gemm_simt(h, PRO_DEC, EPI_TANH_ADD, d));

    GemmArgs j;
    j.M = B; j.N = V; j.K = J;
    j.W = h->out_w; j.bias = h->out_b; j.A = x;
    j.part_val = pval; j.part_idx = pidx; j.part_nan = pnan;
    prof_begin(h);
    if (tc) K2B_TRY(joiner_tc_partials(h, x, nullptr, B, 0, nullptr, nullptr, nullptr, nullptr, pval, pidx, pnan));
    else K2B_TRY(launch_gemm_simt(h, PRO_PLAIN, EPI_ARGMAX, j));
    prof_end(h);

    greedy_select_kernel<<<gb, tb, 0, h->stream>>>(B, nt, pval, pidx, pnan, ctx, tokens, ts, n_out, cap, t, c.blank_id,
                                                   c.unk_id, extra_mask, max_sym, flag,
                                                   (!online && h->lens_active) ? h->lens_dev : nullptr, sub, spf);
    K2B_LAUNCH_CHECK(h);
  }
  if (online) {
    greedy_finish_online_kernel<<<gb, tb, 0, h->stream>>>(B, ctx, hyp_inout);
    K2B_LAUNCH_CHECK(h);
  }
  return K2B_OK;
}

namespace {
// First launch of the fused path, one CTA per hypothesis row: the initial hypothesis state (beam_init_kernel), the context from
// OnlineStream.Hyp for online greedy, and the joiner operand row of frame 0 (joinin_table_kernel) - three launches in one.
__global__ void beam_start_kernel(int B, int K, int V, int J, int blank, const int64_t* __restrict__ hyp, BeamState s0, BeamState s1,
                                  const float* __restrict__ dec_tab, const float* __restrict__ enc, long long enc_stride,
                                  uint8_t* __restrict__ x_img, int* __restrict__ zero, int nzero) {
  const int m = blockIdx.x, b = m / K, slot = m - b * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nzero; i += gridDim.x * blockDim.x) zero[i] = 0;   // the persistent kernel's counters
  int c0 = -1, c1 = blank;
  if (hyp != nullptr && slot == 0) { c0 = (int)hyp[2 * b]; c1 = (int)hyp[2 * b + 1]; }
  if (threadIdx.x == 0) {
    s0.ctx[2 * m] = c0; s0.ctx[2 * m + 1] = c1;
    s1.ctx[2 * m] = -1; s1.ctx[2 * m + 1] = blank;
    s0.lp[m] = slot == 0 ? 0.f : -INFINITY;
    s0.len[m] = 2;
    s0.hash[m] = kHashSeed;
    if (slot == 0) s0.nlive[b] = 1;
  }
  constexpr int kRowTile = 128, kImgTile = 128 * 128;
  const float* erow = enc + (size_t)b * enc_stride;
  const float* drow = dec_tab + ((size_t)(c0 + 1) * V + c1) * J;
  for (int k = threadIdx.x << 3; k < J; k += blockDim.x << 3) {
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(erow + k)), e1 = __ldg(reinterpret_cast<const float4*>(erow + k) + 1);
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(drow + k)), d1 = __ldg(reinterpret_cast<const float4*>(drow + k) + 1);
    const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w}, dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    float x[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {           // tanh(e + d) = 1 - 2 / (1 + exp(2e) * exp(2d)); the table holds exp(2d)
      float r;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(expf(2.f * fminf(fmaxf(ev[i], -21.f), 21.f)), dv[i], 1.f)));
      x[i] = fmaf(-2.f, r, 1.f);
      hi[i] = k2b::ptx::bf16_round(x[i]);
    }
    uint8_t* timg = x_img + ((size_t)(m / kRowTile) * (J / 64) + (k >> 6)) * (2 * kImgTile) + k2b::ptx::sw128_offset(m % kRowTile, k & 63);
    *reinterpret_cast<uint4*>(timg) = make_uint4(k2b::ptx::pack_bf16x2(hi[0], hi[1]), k2b::ptx::pack_bf16x2(hi[2], hi[3]),
                                                 k2b::ptx::pack_bf16x2(hi[4], hi[5]), k2b::ptx::pack_bf16x2(hi[6], hi[7]));
    *reinterpret_cast<uint4*>(timg + kImgTile) =
        make_uint4(k2b::ptx::pack_bf16x2(x[0] - hi[0], x[1] - hi[1]), k2b::ptx::pack_bf16x2(x[2] - hi[2], x[3] - hi[3]),
                   k2b::ptx::pack_bf16x2(x[4] - hi[4], x[5] - hi[5]), k2b::ptx::pack_bf16x2(x[6] - hi[6], x[7] - hi[7]));
  }
}

// beam 1 as online greedy: slot 0's context comes from / goes back to OnlineStream.Hyp (ref OnlineRecognizer.cs:109,125,208)
__global__ void beam_ctx_from_hyp_kernel(int B, const int64_t* __restrict__ hyp, int32_t* __restrict__ ctx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  ctx[2 * b] = (int32_t)hyp[2 * b];
  ctx[2 * b + 1] = (int32_t)hyp[2 * b + 1];
}
__global__ void beam_hyp_from_ctx_kernel(int B, const int32_t* __restrict__ ctx, int64_t* __restrict__ hyp) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  hyp[2 * b] = ctx[2 * b];
  hyp[2 * b + 1] = ctx[2 * b + 1];
}
}  // namespace

bool beam_chunkable(k2b_handle* h, int K) {
  if (h->cfg.precision == K2B_PREC_FP32 || !joiner_tc_supported(h) || h->tab0 == nullptr || !joiner_topk_usable(h, K)) return false;
  bool have = false;
  if (ensure_dec_table(h, &have) != K2B_OK) return false;
  return have && h->opt_unfused_step == 0;
}

// greedy search as beam 1 on the persistent kernel: needs the tcgen05 precisions and the memoised decoder table
bool beam_greedy_usable(k2b_handle* h) {
  if (h->cfg.precision == K2B_PREC_FP32 || !joiner_tc_supported(h) || h->tab0 == nullptr) return false;
  bool have = false;
  if (ensure_dec_table(h, &have) != K2B_OK) return false;
  return have && beam_mega_usable(h, 1);
}

int32_t beam_dev(k2b_handle* h, const float* enc, int B, int T, int K, int64_t* tokens, int32_t* ts, int32_t* n_out,
                 float* score, int cap, int extra_mask, int64_t* hyp_inout, bool greedy, int t0, int Ttot, bool carry,
                 long long enc_stride_in) {
  const k2b_config& c = h->cfg;
  if (Ttot <= 0) { Ttot = T; t0 = 0; }
  if (carry && t0 != 0) return fail(h, K2B_ERR_INVALID, "beam_dev: a carried state starts at the first frame of its chunk");
  const long long enc_stride = enc_stride_in > 0 ? enc_stride_in : (long long)Ttot * c.joiner_dim;   // enc points at frame t0 of a [B,Ttot,J] array
  const bool resume = t0 > 0 || carry, last = !carry && t0 + T >= Ttot;
  const int J = c.joiner_dim, V = c.vocab_size, D = c.decoder_dim;
  const bool tc = c.precision != K2B_PREC_FP32 && joiner_tc_supported(h);   // per-frame tcgen05 joiner (256-column tiles)
  const int N = B * K;
  // greedy callers (no score wanted) get the beam-1 instantiations: top-1 records without sums, one warp per stream
  const int kk = (K == 1 && greedy) ? 1 : (K <= 4 ? 4 : 8);
  // vocabulary tiles per row: the fused joiner may use narrower tiles than the per-frame kernels (same buffer, sized for the larger)
  const int nt_fused = tc ? joiner_topk_tiles(h, N) : 0;
  const int nt = tc ? joiner_tc_tiles(h) : num_vocab_tiles(V);
  K2B_TRY(ensure(h, h->ws_x, sizeof(float) * (size_t)N * J));
  uint8_t* ximg = nullptr;
  bool have_tab = false;         // memoised decoder (fits in HBM up to V ~ 6800 at J = 512): no decoder GEMM in the loop
  // the joiner operand as bf16 hi / lo tile images needs 64-column k-blocks; it is produced from the memoised table (any decoder
  // width) or, when that does not fit, by the tcgen05 decoder (256-column tiles of J)
  if (tc && J % 64 == 0 && h->tab0 != nullptr) K2B_TRY(ensure_dec_table(h, &have_tab));
  if (tc && J % 64 == 0 && (have_tab || decoder_tc_supported(h))) {
    K2B_TRY(ensure(h, h->ws_ximg, joiner_tc_image_bytes(h, N)));
    ximg = static_cast<uint8_t*>(h->ws_ximg.p);
  }
  // per (row, tile): four arrays of the per-frame kernels (8 + 8K bytes) or one record of the fused joiner (beam_partial_words)
  K2B_TRY(ensure(h, h->ws_part, (size_t)N * (size_t)max(nt * (8 + 8 * K), nt_fused * 4 * beam_partial_words(kk))));
  K2B_TRY(ensure(h, h->ws_state, 2 * state_bytes(B, K)));
  K2B_TRY(ensure(h, h->ws_bp, sizeof(int32_t) * (size_t)B * (Ttot > 0 ? Ttot : 1) * K));
  char* p = static_cast<char*>(h->ws_state.p);
  BeamState st[2];
  st[0] = carve_state(p, B, K);
  st[1] = carve_state(p, B, K);
  float* part_m = static_cast<float*>(h->ws_part.p);
  float* part_s = part_m + (size_t)N * nt;
  float* part_tv = part_s + (size_t)N * nt;
  int32_t* part_ti = reinterpret_cast<int32_t*>(part_tv + (size_t)N * nt * K);
  int32_t* bp = static_cast<int32_t*>(h->ws_bp.p);
  float* x = static_cast<float*>(h->ws_x.p);

  const bool greedy_ext = hyp_inout != nullptr;
  if (greedy_ext && K != 1) return fail(h, K2B_ERR_INVALID, "beam_dev: Hyp in / out is a beam-1 (greedy) option");
  const bool unfused = h->opt_unfused_step != 0;
  const bool fused = tc && have_tab && ximg != nullptr && !unfused && joiner_topk_usable(h, K) && T > 0;
  if (t0 > 0 && !fused) return fail(h, K2B_ERR_UNSUPPORTED, "beam_dev: time chunks need the memoised decoder table");
  const bool biased = h->cg_next != nullptr && !greedy;
  if (biased && fused)
    return fail(h, K2B_ERR_UNSUPPORTED, "a context graph (hot words) is served by the cluster kernel (V <= 1024) and by the per-frame merge: "
                                        "k2b_set_option(h, \"unfused_step\", 1) or fp32 precision for this vocabulary");
  if (fused) {
    K2B_TRY(ensure_joiner_assets(h));
    const size_t nsync = beam_mega_sync_ints(h, B, T, K);
    K2B_TRY(ensure(h, h->ws_sync, nsync * sizeof(int)));
    if (!resume) {          // state, Hyp and the operand of frame 0 in one launch
      beam_start_kernel<<<N, 64, 0, h->stream>>>(B, K, V, J, c.blank_id, hyp_inout, st[0], st[1], h->dec_tab, enc, enc_stride, ximg,
                                                static_cast<int*>(h->ws_sync.p), (int)nsync);
      K2B_LAUNCH_CHECK(h);
    } else {                // next time chunk: the state is where the chunk before left it
      K2B_CUDA(h, cudaMemsetAsync(h->ws_sync.p, 0, nsync * sizeof(int), h->stream));
      K2B_TRY(joinin_table_tc(h, st[t0 & 1].ctx, N, enc, enc_stride, K, ximg));
    }
  } else if (!carry) {
    beam_init_kernel<<<(N + 127) / 128, 128, 0, h->stream>>>(B, K, c.blank_id, st[0], st[1]);
    K2B_LAUNCH_CHECK(h);
    if (hyp_inout != nullptr) {
      beam_ctx_from_hyp_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(B, hyp_inout, st[0].ctx);
      K2B_LAUNCH_CHECK(h);
    }
  }
  auto finish = [&](int fin) -> int32_t {
    if (hyp_inout != nullptr) {
      beam_hyp_from_ctx_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(B, st[fin].ctx, hyp_inout);
      K2B_LAUNCH_CHECK(h);
    }
    return beam_backtrace_dev(h, B, K, Ttot, st[fin].lp, st[fin].len, st[fin].nlive, bp, tokens, ts, n_out, score, cap);
  };

  int cur = t0 & 1;
  // memoised decoder + persistent joiner: two launches per frame (joiner, fused merge + next operand), chained by programmatic
  // dependent launches. K2B_UNFUSED_STEP=1 keeps the three-launch sequence below (comparison runs).
  if (fused) {
    if (beam_mega_usable(h, K) && !(h->profile_on && h->prof_which != 0)) {       // the whole time loop in one launch
      BeamStatePtrs sp[2];
      for (int i = 0; i < 2; ++i)
        sp[i] = BeamStatePtrs{st[i].ctx, st[i].lp, st[i].len, reinterpret_cast<unsigned long long*>(st[i].hash), st[i].nlive};
      prof_begin(h);
      const GreedyOutPtrs go{tokens, ts, n_out, hyp_inout, cap};
      const int32_t ms = beam_mega_tc(h, enc, B, T, K, ximg, part_m, sp[0], sp[1], bp,
                                      h->lens_active ? h->lens_dev : nullptr, extra_mask, kk, kk == 1 ? &go : nullptr, true, t0, Ttot,
                                      enc_stride);
      prof_end(h);
      if (ms != kMegaUnavailable) {
        K2B_TRY(ms);
        if (kk == 1 || !last) return K2B_OK;   // greedy: the merge warps have written tokens, timestamps, counts and Hyp
        return finish(Ttot & 1);
      }
    }
    for (int t = 0; t < T; ++t) {
      if (h->prof_which == 0) prof_begin(h);
      K2B_TRY(joiner_topk_tc(h, ximg, N, K, kk, part_m));
      if (h->prof_which == 0) prof_end(h);
      if (h->prof_which == 2) prof_begin(h);
      K2B_CUDA(h, launch_pdl(kk == 1 ? beam_step_kernel<1> : kk == 4 ? beam_step_kernel<4> : beam_step_kernel<8>, dim3(B), dim3(128), 0,
                              h->stream, B, K, V, nt_fused, Ttot, t0 + t, (int)c.blank_id, (int)c.unk_id, extra_mask,
                              (const float*)part_m, st[cur], st[cur ^ 1], bp, (const int32_t*)(h->lens_active ? h->lens_dev : nullptr), (const float*)h->dec_tab,
                              (const float*)(t + 1 < T ? enc + (size_t)(t + 1) * J : nullptr), enc_stride, J, ximg,
                              (long long*)(h->timeline != nullptr ? h->timeline + (size_t)(h->timeline_frame % 64) * 148 * 8 : nullptr)));
      K2B_LAUNCH_CHECK(h);
      h->timeline_frame++;
      if (h->prof_which == 2) prof_end(h);
      cur ^= 1;
    }
    return last ? finish(cur) : K2B_OK;
  }
  if (greedy_ext) return fail(h, K2B_ERR_UNSUPPORTED, "beam_dev: greedy options need the memoised decoder table");
  h->ll_clean_ptr = nullptr;             // (the persistent greedy kernel's tagged records share the partials buffer)
  if (enc_stride_in > 0 && enc_stride_in != (long long)T * J) return fail(h, K2B_ERR_UNSUPPORTED, "beam_dev: strided frames need the fused path");
  for (int t = 0; t < T; ++t) {
    GemmArgs d;
    d.M = N; d.N = J; d.K = D;
    d.W = h->dec_w; d.bias = h->dec_b;
    d.ctx = st[cur].ctx; d.tab0 = h->tab0; d.tab1 = h->tab1; d.V = V; d.neg_wrap = c.neg_id_mode == K2B_NEGID_WRAP;
    d.blank = c.blank_id;
    d.enc = enc + (size_t)t * J; d.enc_stride = (long long)T * J; d.rows_per_stream = K;
    d.C = x;
    // tcgen05 decoder: its epilogue leaves x = tanh(enc + dec) as bf16 hi/lo tile images, which the joiner's loader warp
    // fetches by TMA (no fp32 round trip, no per-CTA conversion)
    const bool img = ximg != nullptr;
    if (h->prof_which == 1) prof_begin(h);
    if (img && have_tab) K2B_TRY(joinin_table_tc(h, st[cur].ctx, N, enc + (size_t)t * J, (long long)T * J, K, ximg));
    else if (img) K2B_TRY(decoder_joinin_tc(h, st[cur].ctx, N, enc + (size_t)t * J, (long long)T * J, K, nullptr, ximg));
    else K2B_TRY(launch_gemm_simt(h, PRO_DEC, EPI_TANH_ADD, d));
    if (h->prof_which == 1) prof_end(h);

    GemmArgs j;
    j.M = N; j.N = V; j.K = J;
    j.W = h->out_w; j.bias = h->out_b; j.A = x;
    j.part_m = part_m; j.part_s = part_s; j.part_tv = part_tv; j.part_ti = part_ti; j.topk = K;
    if (h->prof_which == 0) prof_begin(h);
    if (tc) K2B_TRY(joiner_tc_partials(h, x, img ? ximg : nullptr, N, K, part_m, part_s, part_tv, part_ti, nullptr, nullptr, nullptr));
    else K2B_TRY(launch_gemm_simt(h, PRO_PLAIN, EPI_TOPK, j));
    if (h->prof_which == 0) prof_end(h);

    if (h->prof_which == 2) prof_begin(h);
    K2B_CUDA(h, launch_pdl(beam_select_kernel, dim3((B + 3) / 4), dim3(128), 0, h->stream, B, K, V, nt, T, t, (int)c.blank_id,
                            (int)c.unk_id, extra_mask, (const float*)part_m, (const float*)part_s, (const float*)part_tv,
                            (const int32_t*)part_ti, st[cur], st[cur ^ 1], bp,
                            (const int32_t*)(h->lens_active ? h->lens_dev : nullptr),
                            (const int32_t*)(biased ? h->cg_next : nullptr), (const float*)(biased ? h->cg_delta : nullptr)));
    K2B_LAUNCH_CHECK(h);
    if (h->prof_which == 2) prof_end(h);
    cur ^= 1;
  }
  if (carry) return K2B_OK;
  return beam_backtrace_dev(h, B, K, T, st[cur].lp, st[cur].len, st[cur].nlive, bp, tokens, ts, n_out, score, cap,
                            biased ? st[cur].cst : nullptr);
}

int32_t beam_backtrace_dev(k2b_handle* h, int B, int K, int T, const float* lp, const int32_t* len, const int32_t* nlive,
                           const int32_t* bp, int64_t* tokens, int32_t* ts, int32_t* n_out, float* score, int cap, const int32_t* cst) {
  BeamState fin{};
  fin.lp = const_cast<float*>(lp);
  fin.len = const_cast<int32_t*>(len);
  fin.nlive = const_cast<int32_t*>(nlive);
  fin.cst = const_cast<int32_t*>(cst);
  const size_t smem = sizeof(int32_t) * ((size_t)T * K + 2 * (size_t)T);
  if (smem > 200 * 1024) return fail(h, K2B_ERR_UNSUPPORTED, "modified_beam_search: T*K too large for the back-trace");
  if (smem > 48 * 1024)
    K2B_CUDA(h, cudaFuncSetAttribute(beam_backtrace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  beam_backtrace_kernel<<<B, 32, smem, h->stream>>>(B, K, T, fin, bp, tokens, ts, n_out, score, cap,
                                                    cst != nullptr ? h->cg_resid : nullptr);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

}  // namespace k2b
