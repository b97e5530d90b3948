// Hypothesis state of modified_beam_search and the per-stream frame step ("hyp_merge" + the next frame's joiner operand), shared by
// the stand-alone merge kernel (search.cu) and the persistent beam-search kernel (joiner_tc.cu).
#pragma once
#include <math.h>
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace k2b {

struct BeamState {
  int32_t* ctx;    // [N,2]
  float* lp;       // [N]
  int32_t* len;    // [N]
  uint64_t* hash;  // [N]
  int32_t* nlive;  // [B]
  int32_t* cst = nullptr;   // [N] context-graph (hot word) state; only the per-frame fp32 merge and the cluster kernel bias
};

constexpr uint64_t kHashSeed = 0x9E3779B97F4A7C15ull;

__device__ __forceinline__ uint64_t hash_push(uint64_t h, int tok) {
  h = (h ^ (uint64_t)(uint32_t)(tok + 1)) * 0x100000001B3ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h;
}

__device__ __forceinline__ float logaddexp_f(float a, float b) {
  const float mx = fmaxf(a, b), mn = fminf(a, b);
  if (mx == -INFINITY) return -INFINITY;
  return mx + log1pf(expf(mn - mx));
}

// Partial record of one (row, vocabulary tile) as the fused joiner writes it: (max, sum-exp, KB values, KB indices), padded to whole
// 16-byte vectors: three vector stores per row instead of ten scalar ones (the release that publishes a tile waits for them)
template <int KB>
constexpr int kBeamRecWords = (2 + 2 * KB + 3) & ~3;

// order-preserving integer image of a float
__device__ __forceinline__ int fkey_s(float f) { const int k = __float_as_int(f); return k ^ ((k >> 31) & 0x7fffffff); }
__device__ __forceinline__ float funkey_s(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

// One frame step of stream s by a group of 128 threads (tid 0..127, synchronised with named barrier `bar_id`):
//   A. warp h <-> live hypothesis h: log_softmax constants from its tile partials, its K best extensions (all loads of a
//      hypothesis are issued before the first use: the partials sit in L2, the step pays latency, not bandwidth);
//   B. warp 0: the stream's top K over the K*K survivors (value desc, flat index desc), extension, dedupe by token-sequence hash
//      with log-add in rank order, compaction, back-pointer record;
//   C. all threads: x[m,:] = tanh(enc[s,t+1] + decoder(ctx[m])) of the K new hypotheses from the memoised decoder table, written
//      as the bf16 hi / lo tile images the joiner's loader warp fetches.
// part_rec [N, nt, kBeamRecWords<KB>] is what joiner_topk_kernel<KB> writes.
// Everything another CTA may have produced during the same launch (partials, state) is read with ld.global.cg.
// mask3: a third non-emitting id (the literal 1 of ref OnlineRecognizer.cs:181), or -1.
// KB = 1, 4 or 8: compile-time bound of the beam (the step is latency-bound and most of its instructions are executed once, so its
// code size is its run time: loops are unrolled to exactly KB levels). c_v / c_f [KB*KB] and s_ctx [2*KB] are shared scratch.
// the next frame's encoder values of thread tid's first operand item (8 consecutive k of hypothesis tid / (J/8)): they do not depend
// on the search, so the callers fetch them before they wait
__device__ __forceinline__ void beam_merge_prefetch(int tid, int s, int K, int J, const float* enc_next, long long enc_stride, float4* pe0,
                                                    float4* pe1) {
  *pe0 = *pe1 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int per_row = J >> 3;
  if (enc_next != nullptr && tid < K * per_row) {
    const float* erow = enc_next + (size_t)s * enc_stride + ((tid % per_row) << 3);
    *pe0 = __ldg(reinterpret_cast<const float4*>(erow));
    *pe1 = __ldg(reinterpret_cast<const float4*>(erow) + 1);
  }
}

template <int KB>
__device__ __forceinline__ void beam_merge_stream(
    int tid, int bar_id, int s, int K, int V, int nt, int T, int t, int blank, int unk, int mask3, const float* part_rec,
    const BeamState& in, const BeamState& out, int32_t* bp, const int32_t* lens,
    const float* dec_tab, const float* enc_next, long long enc_stride, int J, uint8_t* x_img, float4 pe0, float4 pe1, float* c_v,
    int* c_f,
    int* s_ctx, long long* tl, uint32_t tag = 0u, int* abort_flag = nullptr, int* bad = nullptr) {
  // tag != 0 (persistent kernel, nt <= 64, V < 65535): TAGGED records - every 16-byte vector of a record is (three data words, the
  // epoch tag of its frame), the data words being max, sum-exp, KB values and the KB indices packed two 16-bit halves to a word;
  // the joiner CTAs announce nothing, every lane polls its own records until all their vectors carry the expected tag (see
  // greedy_merge_warp). A time-out / abort sets *bad (shared) and the records read as empty.
  constexpr int kNone = (int)0x80000000;
  // per-lane register batch: the records of tiles lane and lane + 32 (nt <= 64 without a tail pass), KB candidates each
  constexpr int kPairs = 2, kCands = kPairs * KB, RW = kBeamRecWords<KB>;
  const unsigned full = 0xffffffffu;
  const int warp = tid >> 5, lane = tid & 31;
  const bool build = enc_next != nullptr;
  const int nl = __ldcg(in.nlive + s);
  const bool frozen = lens != nullptr && t >= __ldg(lens + s);      // ragged batch: past the end of this stream
  if (tid < KB * KB) { c_v[tid] = -INFINITY; c_f[tid] = -1; }
  // parents' state, lane q of warp 0 <-> hypothesis q
  uint64_t p_hash = kHashSeed;
  int p_len = 2, p_c0 = -1, p_c1 = blank;
  float p_lp = -INFINITY;
  if (warp == 0 && lane < K) {
    const size_t o = (size_t)s * K + lane;
    p_hash = __ldcg(in.hash + o); p_len = __ldcg(in.len + o); p_c0 = __ldcg(in.ctx + 2 * o); p_c1 = __ldcg(in.ctx + 2 * o + 1);
    p_lp = __ldcg(in.lp + o);
  }
  k2b::ptx::named_bar_sync(bar_id, 128);
  if (frozen) {
    if (warp == 0) {
      if (lane < K) {
        const size_t o = (size_t)s * K + lane;
        out.ctx[2 * o] = p_c0; out.ctx[2 * o + 1] = p_c1;
        out.lp[o] = p_lp; out.len[o] = p_len; out.hash[o] = p_hash;
        bp[((size_t)s * T + t) * K + lane] = lane < nl ? (lane << 28) : 0;
        s_ctx[2 * lane] = p_c0; s_ctx[2 * lane + 1] = p_c1;
      }
      if (lane == 0) out.nlive[s] = nl;
    }
  } else {
    if (tag != 0u) {
      // tagged records: ONE warp polls (the last vector of the records of hypothesis 0, with a pause between rounds - 128 threads
      // polling every vector of every record cost 10 % on cfg4, whose weight stream runs at the L2 -> SM limit); the records of
      // the stream's other rows come from the same joiner CTAs a few cycles apart, and every load below still verifies its tags
      if (warp == 0) {
        const float* r0 = part_rec + (size_t)s * K * nt * RW;
        const long long c0 = clock64();
        int spins = 0;
#pragma unroll 1
        while (true) {
          bool all = true;
#pragma unroll
          for (int u = 0; u < kPairs; ++u) {
            const int i = lane + 32 * u;
            if (i < nt) {
              uint32_t tg;
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(tg) : "l"(r0 + (size_t)i * RW + RW - 1) : "memory");
              all &= tg == tag;
            }
          }
          if (__all_sync(full, all)) break;
          __nanosleep(20);
          if (((++spins) & 15) == 0) {
            int ab;
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(ab) : "l"(abort_flag) : "memory");
            if (ab != 0 || clock64() - c0 > k2b::ptx::kWaitTimeoutCycles) { if (lane == 0) *bad = 1; break; }
          }
        }
      }
      k2b::ptx::named_bar_sync(bar_id, 128);
    }
    // ---- A: per live hypothesis --------------------------------------------------------------------------------------------
#pragma unroll 1
    for (int h = warp; h < K; h += 4) {        // the loads of row h are issued before the live count has arrived (dead rows are
                                               // valid memory: the joiner reduced them like any other row)
      const size_t row = (size_t)s * K + h;
      const float* rrow = part_rec + row * nt * RW;          // this row's records: [nt][RW] = (max, sum-exp, KB values, KB indices)
      float pm[kPairs], ps[kPairs], cval[kCands];
      int cidx[kCands];
#pragma unroll
      for (int u = 0; u < kPairs; ++u) {
        const int i = lane + 32 * u;
        uint32_t w[RW];
        if (tag == 0u) {
#pragma unroll
          for (int v = 0; v < RW / 4; ++v) {
            uint4 q4 = make_uint4(0xff800000u, 0u, 0xff800000u, 0xff800000u);
            if (i < nt) q4 = __ldcg(reinterpret_cast<const uint4*>(rrow + (size_t)i * RW) + v);
            w[4 * v] = q4.x; w[4 * v + 1] = q4.y; w[4 * v + 2] = q4.z; w[4 * v + 3] = q4.w;
          }
        } else {
          // plain layout out of the tagged one: data word j sits in vector j / 3 at position j % 3
          uint32_t d[3 * (RW / 4)];
#pragma unroll
          for (int j = 0; j < 3 * (RW / 4); ++j) d[j] = (j == 0) ? 0xff800000u : 0u;
          if (i < nt) {
            const uint4* rp = reinterpret_cast<const uint4*>(rrow + (size_t)i * RW);
            const long long c0 = clock64();
            int spins = 0;
#pragma unroll 1
            while (true) {
              bool all = true;
#pragma unroll
              for (int v = 0; v < RW / 4; ++v) {
                uint4 q4;
                asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q4.x), "=r"(q4.y), "=r"(q4.z), "=r"(q4.w) : "l"(rp + v) : "memory");
                d[3 * v] = q4.x; d[3 * v + 1] = q4.y; d[3 * v + 2] = q4.z;
                all &= q4.w == tag;
              }
              if (all) break;
              if (((++spins) & 63) == 0) {
                int ab;
                asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(ab) : "l"(abort_flag) : "memory");
                if (ab != 0 || clock64() - c0 > k2b::ptx::kWaitTimeoutCycles) {
                  *bad = 1;
#pragma unroll
                  for (int j = 0; j < 3 * (RW / 4); ++j) d[j] = (j == 0) ? 0xff800000u : 0u;
                  d[2 + KB] = 0xffffffffu;
                  break;
                }
              }
            }
          }
          w[0] = d[0]; w[1] = d[1];
#pragma unroll
          for (int j = 0; j < KB; ++j) {
            w[2 + j] = d[2 + j];
            const uint32_t h16 = (d[2 + KB + (j >> 1)] >> (16 * (j & 1))) & 0xffffu;
            w[2 + KB + j] = h16 == 0xffffu ? 0xffffffffu : h16;
          }
        }
        pm[u] = i < nt ? __uint_as_float(w[0]) : -INFINITY;
        ps[u] = i < nt ? __uint_as_float(w[1]) : 0.f;
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          cval[u * KB + j] = __uint_as_float(w[2 + j]);
          cidx[u * KB + j] = i < nt ? (int)w[2 + KB + j] : -1;
        }
      }
      const float lp = __ldcg(in.lp + row);
      if (h >= nl) continue;
      int mk = max(fkey_s(pm[0]), fkey_s(pm[1]));           // absent tiles read as -inf
#pragma unroll 1
      for (int i = lane + 32 * kPairs; i < nt; i += 32) mk = max(mk, fkey_s(__ldcg(rrow + (size_t)i * RW)));
      const float mx = funkey_s(__reduce_max_sync(full, mk));
      float sum = 0.f;
#pragma unroll
      for (int u = 0; u < kPairs; ++u) sum += (pm[u] > -INFINITY) ? ps[u] * __expf(pm[u] - mx) : 0.f;
#pragma unroll 1
      for (int i = lane + 32 * kPairs; i < nt; i += 32) {
        const float m2 = __ldcg(rrow + (size_t)i * RW);
        sum += (m2 > -INFINITY) ? __ldcg(rrow + (size_t)i * RW + 1) * __expf(m2 - mx) : 0.f;
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(full, sum, o);
      const float ls = __logf(sum);
      if (tl != nullptr && tid == 0) tl[0] = clock64();   // diagnostic stamp
      // this lane's candidates as (key, flat index); same operation order as log_softmax(x) + lp : ((x - max) - log(sum)) + lp
      int ck[kCands], cf[kCands];
#pragma unroll
      for (int u = 0; u < kCands; ++u) {
        const float v = ((cval[u] - mx) - ls) + lp;
        const bool okc = (cidx[u] >= 0) & (v == v);
        ck[u] = okc ? fkey_s(v) : kNone;
        cf[u] = okc ? h * V + cidx[u] : -1;
      }
      // more than 64 tiles (rare): a lane keeps its best kCands >= K candidates (its worst one is replaced)
#pragma unroll 1
      for (int i = lane + 32 * kPairs; i < nt; i += 32) {
#pragma unroll 1
        for (int j = 0; j < KB; ++j) {
          const int idx = (int)__float_as_uint(__ldcg(rrow + (size_t)i * RW + 2 + KB + j));
          const float v = ((__ldcg(rrow + (size_t)i * RW + 2 + j) - mx) - ls) + lp;
          const bool okc = (idx >= 0) & (v == v);
          const int key = okc ? fkey_s(v) : kNone, f = okc ? h * V + idx : -1;
          int wk = ck[0], wf = cf[0];
#pragma unroll
          for (int u = 1; u < kCands; ++u) {
            const bool lower = (ck[u] < wk) | ((ck[u] == wk) & (cf[u] < wf));
            wk = lower ? ck[u] : wk; wf = lower ? cf[u] : wf;
          }
          const bool take = (key > wk) | ((key == wk) & (f > wf));
          bool done = !take;
#pragma unroll
          for (int u = 0; u < kCands; ++u) {
            const bool hit = !done & (ck[u] == wk) & (cf[u] == wf);
            ck[u] = hit ? key : ck[u]; cf[u] = hit ? f : cf[u];
            done |= hit;
          }
        }
      }
      if (tl != nullptr && tid == 0) tl[1] = clock64();   // diagnostic stamp
      // K rounds of warp arg-best over all registers: REDUX on the key, then on the flat index among the ties
#pragma unroll
      for (int r = 0; r < KB; ++r) {
        if (r < K) {
          int lk = ck[0];
#pragma unroll
          for (int u = 1; u < kCands; ++u) lk = max(lk, ck[u]);
          const int wk = __reduce_max_sync(full, lk);
          int lf = -1;
#pragma unroll
          for (int u = 0; u < kCands; ++u) lf = max(lf, (ck[u] == wk) ? cf[u] : -1);
          const int wf = __reduce_max_sync(full, lf);      // flat indices are unique: exactly one register of one lane matches
#pragma unroll
          for (int u = 0; u < kCands; ++u) ck[u] = ((ck[u] == wk) & (cf[u] == wf)) ? kNone : ck[u];
          if (lane == r) { c_v[h * K + r] = wf >= 0 ? funkey_s(wk) : -INFINITY; c_f[h * K + r] = wf; }
        }
      }
    }
    k2b::ptx::named_bar_sync(bar_id, 128);
    // ---- B: the stream's top K, extension, merge ---------------------------------------------------------------------------
    if (tl != nullptr && tid == 0) tl[2] = clock64();   // diagnostic stamp: phase A done (barrier passed)
    if (warp == 0) {
      int tk[2], tf[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = lane + 32 * u;
        const bool okc = i < K * K && c_f[i < KB * KB ? i : 0] >= 0;
        tk[u] = okc ? fkey_s(c_v[i < KB * KB ? i : 0]) : kNone;
        tf[u] = okc ? c_f[i < KB * KB ? i : 0] : -1;
      }
      float my_v = -INFINITY;
      int my_f = -1;
      if constexpr (KB * KB <= 32) {
        // one candidate per lane: its rank is the number of better candidates, found with independent shuffles (no chain of
        // reductions); the K best reach lanes 0..K-1 through shared memory (valid candidates have distinct flat indices)
        const int mk = tk[0], mf = tf[0];
        int rank = 0;
#pragma unroll
        for (int j = 0; j < KB * KB; ++j) {
          const int ok = __shfl_sync(full, mk, j), of = __shfl_sync(full, mf, j);
          rank += ((ok > mk) | ((ok == mk) & (of > mf))) ? 1 : 0;
        }
        __syncwarp();
        if (lane < KB) c_f[lane] = -1;
        __syncwarp();
        if (mf >= 0 && rank < K) { c_v[rank] = funkey_s(mk); c_f[rank] = mf; }
        __syncwarp();
        if (lane < K) { my_f = c_f[lane]; my_v = my_f >= 0 ? c_v[lane] : -INFINITY; }
      } else {
#pragma unroll
        for (int r = 0; r < KB; ++r) {
          if (r < K) {
            const int wk = __reduce_max_sync(full, max(tk[0], tk[1]));
            const int wf = __reduce_max_sync(full, max((tk[0] == wk) ? tf[0] : -1, (tk[1] == wk) ? tf[1] : -1));
            tk[0] = ((tk[0] == wk) & (tf[0] == wf)) ? kNone : tk[0];
            tk[1] = ((tk[1] == wk) & (tf[1] == wf)) ? kNone : tk[1];
            if (lane == r) { my_v = wf >= 0 ? funkey_s(wk) : -INFINITY; my_f = wf; }
          }
        }
      }
      if (tl != nullptr && tid == 0) tl[3] = clock64();   // diagnostic stamp
      // lane r < K: the r-th extension in rank order
      const bool cand = lane < K && my_f >= 0;
      const int par = cand ? my_f / V : 0;
      const uint64_t ph = __shfl_sync(full, p_hash, par);
      const int pl = __shfl_sync(full, p_len, par), pc0 = __shfl_sync(full, p_c0, par), pc1 = __shfl_sync(full, p_c1, par);
      int tok = -1, c0 = -1, c1 = blank, ln = 2;
      uint64_t hs = kHashSeed;
      if (cand) {
        const int y = my_f - par * V;
        hs = ph; ln = pl; c0 = pc0; c1 = pc1;
        if (y != blank && y != unk && y != mask3) {      // ys unchanged for blank / unk (/ the literal 1 of the online loop)
          tok = y;
          hs = hash_push(hs, y);
          ln += 1;
          c0 = c1;
          c1 = y;
        }
      }
      if (tl != nullptr && tid == 0) tl[4] = clock64();   // diagnostic stamp
      // dedupe: first earlier lane holding the same token sequence; log-add the merged scores into their root in rank order
      int root = lane;
      float lp = my_v;
#pragma unroll
      for (int q = 0; q < KB; ++q) {
        if (q < K) {
          const uint64_t qh = __shfl_sync(full, hs, q);
          const int ql = __shfl_sync(full, ln, q);
          const int q0 = __shfl_sync(full, c0, q);
          const int q1 = __shfl_sync(full, c1, q);
          const int qc = __shfl_sync(full, (int)cand, q);
          if (cand && qc && q < lane && root == lane && qh == hs && ql == ln && q0 == c0 && q1 == c1) root = q;
        }
      }
#pragma unroll 1
      for (int q = 0; q < K; ++q) {
        const int qroot = __shfl_sync(full, root, q);
        const float qv = __shfl_sync(full, my_v, q);
        const int qc = __shfl_sync(full, (int)cand, q);
        if (cand && qc && q != lane && qroot == lane) lp = logaddexp_f(lp, qv);
      }
      if (tl != nullptr && tid == 0) tl[5] = clock64();   // diagnostic stamp
      const bool is_root = cand && root == lane;
      const unsigned roots = __ballot_sync(full, is_root);
      const int nnew = __popc(roots);
      if (is_root) {
        const int slot = __popc(roots & ((1u << lane) - 1u));
        const size_t o = (size_t)s * K + slot;
        out.ctx[2 * o] = c0;
        out.ctx[2 * o + 1] = c1;
        out.lp[o] = lp;
        out.len[o] = ln;
        out.hash[o] = hs;
        bp[((size_t)s * T + t) * K + slot] = (par << 28) | (tok + 1);
        s_ctx[2 * slot] = c0; s_ctx[2 * slot + 1] = c1;
      }
      if (lane >= nnew && lane < K) {      // dead slots keep a valid context for the next joiner operand
        const size_t o = (size_t)s * K + lane;
        out.ctx[2 * o] = -1;
        out.ctx[2 * o + 1] = blank;
        out.lp[o] = -INFINITY;
        out.len[o] = 2;
        out.hash[o] = kHashSeed;
        bp[((size_t)s * T + t) * K + lane] = 0;
        s_ctx[2 * lane] = -1; s_ctx[2 * lane + 1] = blank;
      }
      if (lane == 0) out.nlive[s] = nnew;
    }
  }
  k2b::ptx::named_bar_sync(bar_id, 128);
  if (tl != nullptr && tid == 0) atomicMax(reinterpret_cast<unsigned long long*>(tl + 6), (unsigned long long)clock64());
  // ---- C: the next frame's joiner operand of this stream's K hypotheses -------------------------------------------------------
  if (!build) return;
  constexpr int kRowTile = 128, kImgTile = 128 * 128;        // rows per image tile, bytes of one 128 x 64 bf16 tile
  const int per_row = J >> 3;                                // work item = 8 consecutive k of one hypothesis: 16-byte image stores
  for (int it = tid; it < K * per_row; it += 128) {
    const int q = it / per_row, k = (it - q * per_row) << 3;
    const float* erow = enc_next + (size_t)s * enc_stride + k;
    const float* drow = dec_tab + ((size_t)(s_ctx[2 * q] + 1) * V + s_ctx[2 * q + 1]) * J + k;
    // (the caller fetched the frame values of this thread's first item before it waited for the partials)
    const float4 e0 = it == tid ? pe0 : __ldg(reinterpret_cast<const float4*>(erow));
    const float4 e1 = it == tid ? pe1 : __ldg(reinterpret_cast<const float4*>(erow) + 1);
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(drow)), d1 = __ldg(reinterpret_cast<const float4*>(drow) + 1);
    const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w}, dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    float x[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {           // tanh(e + d) = 1 - 2 / (1 + exp(2e) * exp(2d)); the table holds exp(2d)
      float r;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(expf(2.f * fminf(fmaxf(ev[i], -21.f), 21.f)), dv[i], 1.f)));
      x[i] = fmaf(-2.f, r, 1.f);
      hi[i] = k2b::ptx::bf16_round(x[i]);
    }
    const int m = s * K + q;
    uint8_t* timg = x_img + ((size_t)(m / kRowTile) * (J / 64) + (k >> 6)) * (2 * kImgTile) + k2b::ptx::sw128_offset(m % kRowTile, k & 63);
    *reinterpret_cast<uint4*>(timg) = make_uint4(k2b::ptx::pack_bf16x2(hi[0], hi[1]), k2b::ptx::pack_bf16x2(hi[2], hi[3]),
                                                 k2b::ptx::pack_bf16x2(hi[4], hi[5]), k2b::ptx::pack_bf16x2(hi[6], hi[7]));
    *reinterpret_cast<uint4*>(timg + kImgTile) =
        make_uint4(k2b::ptx::pack_bf16x2(x[0] - hi[0], x[1] - hi[1]), k2b::ptx::pack_bf16x2(x[2] - hi[2], x[3] - hi[3]),
                   k2b::ptx::pack_bf16x2(x[4] - hi[4], x[5] - hi[5]), k2b::ptx::pack_bf16x2(x[6] - hi[6], x[7] - hi[7]));
  }
  if (tl != nullptr && tid == 0) atomicMax(reinterpret_cast<unsigned long long*>(tl + 7), (unsigned long long)clock64());
}



// Greedy search (beam 1) frame step of stream s by ONE warp: the arg-best over the tile records of its row (value desc, index
// desc = the reference's fold, ref OfflineRecognizer.cs:145-159), the emission test, context shift, back-pointer, and the next
// frame's joiner operand row. With a single hypothesis the log-softmax is a constant shift of every candidate, so neither the
// maxima nor the sums are read and the score stays 0; nothing goes through shared memory, so the four merge warps of a CTA step
// four streams at once. part_rec [B, nt, kBeamRecWords<1>] as joiner_topk_kernel<1> writes it. The emitted tokens / timestamps go
// straight to the caller's arrays (the hypothesis length counts them), so beam 1 needs neither back-pointers nor a back-trace
// launch; the last frame also leaves the count and, for online chunks, the context (OnlineStream.Hyp, ref OnlineRecognizer.cs:208).
struct GreedyOut {
  int64_t* tokens;   // [B,cap]
  int32_t* ts;       // [B,cap]
  int32_t* n;        // [B]
  int64_t* hyp;      // [B,2] or null
  int cap;
};
// `tag` != 0 (persistent kernel): the joiner CTAs do not announce their records through a counter - word 0 of a record is the
// epoch tag of the frame it belongs to, written with the other three words by ONE 16-byte store, and every lane polls its own
// records until the tag is the expected one: the producer needs no barrier and no gpu-scope release (a MEMBAR waiting for the
// CTA's outstanding stores), the consumer no separate round trip for the counter. Returns false on a time-out / abort.
__device__ __forceinline__ bool greedy_merge_warp(int lane, int s, int V, int nt, int T, int t, int blank, int unk, int mask3,
                                                  const float* part_rec, const BeamState& in, const BeamState& out, const GreedyOut& go,
                                                  const int32_t* lens, const float* dec_tab, const float* enc_next,
                                                  long long enc_stride, int J, uint8_t* x_img, uint32_t tag = 0u,
                                                  int* abort_flag = nullptr) {
  constexpr int kNone = (int)0x80000000, RW = kBeamRecWords<1>;
  const unsigned full = 0xffffffffu;
  const float* rrow = part_rec + (size_t)s * nt * RW;
  // the stream's own state (written by this warp one frame ago): fetched BEFORE the wait for the records, not behind it
  int c0 = __ldcg(in.ctx + 2 * s), c1 = __ldcg(in.ctx + 2 * s + 1), ln = __ldcg(in.len + s);
  const bool frozen = lens != nullptr && t >= __ldg(lens + s);
  int bk = kNone, bf = -1;
  bool good = true;
  for (int i = lane; i < nt; i += 32) {
    uint4 q;
    if (tag == 0u) {
      q = __ldcg(reinterpret_cast<const uint4*>(rrow + (size_t)i * RW));
    } else {
      const uint4* rp = reinterpret_cast<const uint4*>(rrow + (size_t)i * RW);
      asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(rp) : "memory");
      if (q.x != tag) {
        const long long c0 = clock64();
        int spins = 0;
#pragma unroll 1
        while (true) {
          asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(rp) : "memory");
          if (q.x == tag) break;
          if (((++spins) & 63) == 0) {
            int ab;
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(ab) : "l"(abort_flag) : "memory");
            if (ab != 0 || clock64() - c0 > ptx::kWaitTimeoutCycles) { good = false; break; }
          }
        }
      }
    }
    const float v = __uint_as_float(q.z);
    const int idx = (int)q.w;
    const bool okc = (idx >= 0) & (v == v);
    const int key = okc ? fkey_s(v) : kNone, f = okc ? idx : -1;
    const bool better = (key > bk) | ((key == bk) & (f > bf));
    bk = better ? key : bk; bf = better ? f : bf;
  }
  if (tag != 0u && !__all_sync(full, good)) return false;
  const int wk = __reduce_max_sync(full, bk);
  const int y = __reduce_max_sync(full, bk == wk ? bf : -1);
  const bool emit = !frozen && y >= 0 && y != blank && y != unk && y != mask3;
  if (lane == 0) {
    if (emit) {
      const int pos = ln - 2;                 // the length starts at the two context tokens
      if (pos < go.cap) { go.tokens[(size_t)s * go.cap + pos] = y; go.ts[(size_t)s * go.cap + pos] = t; }
    }
  }
  if (emit) { c0 = c1; c1 = y; ln += 1; }
  if (lane == 0) {
    out.ctx[2 * s] = c0; out.ctx[2 * s + 1] = c1;
    out.lp[s] = 0.f; out.len[s] = ln; out.hash[s] = kHashSeed; out.nlive[s] = 1;
    if (t == T - 1) {
      go.n[s] = ln - 2;
      if (go.hyp != nullptr) { go.hyp[2 * s] = c0; go.hyp[2 * s + 1] = c1; }
    }
  }
  if (enc_next == nullptr) return true;
  constexpr int kRowTile = 128, kImgTile = 128 * 128;
  const float* erow = enc_next + (size_t)s * enc_stride;
  const float* drow = dec_tab + ((size_t)(c0 + 1) * V + c1) * J;
  for (int k = lane << 3; k < J; k += 256) {
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
    uint8_t* timg = x_img + ((size_t)(s / kRowTile) * (J / 64) + (k >> 6)) * (2 * kImgTile) + k2b::ptx::sw128_offset(s % kRowTile, k & 63);
    *reinterpret_cast<uint4*>(timg) = make_uint4(k2b::ptx::pack_bf16x2(hi[0], hi[1]), k2b::ptx::pack_bf16x2(hi[2], hi[3]),
                                                 k2b::ptx::pack_bf16x2(hi[4], hi[5]), k2b::ptx::pack_bf16x2(hi[6], hi[7]));
    *reinterpret_cast<uint4*>(timg + kImgTile) =
        make_uint4(k2b::ptx::pack_bf16x2(x[0] - hi[0], x[1] - hi[1]), k2b::ptx::pack_bf16x2(x[2] - hi[2], x[3] - hi[3]),
                   k2b::ptx::pack_bf16x2(x[4] - hi[4], x[5] - hi[5]), k2b::ptx::pack_bf16x2(x[6] - hi[6], x[7] - hi[7]));
  }
  return true;
}

}  // namespace k2b
