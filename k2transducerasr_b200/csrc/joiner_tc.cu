// Per-frame fused joiner for large vocabularies (cfg4: V = 5537): logits tile = x * out_w^T + out_b on tcgen05, reduced to
// softmax / top-k partials in the epilogue - the logits never reach HBM (ref JoinerProj, OfflineProjOfTransducer.cs:125-152,
// followed by the log_softmax / top-k of modified_beam_search [EXT]).
//
// Persistent: one CTA per SM walks the 128 x 160 output tiles (280 at M = 1024, V = 5537: two per CTA) with
//   warp 0      loader: bulk TMA copies of the pre-swizzled bf16 hi / lo images of both operands (x was left in that form by the
//               operand kernel, out_w is packed at load time) into a 2-stage ring of 72 KB stages;
//   warp 1      MMA issuer (warp-uniform loop, one elected lane): tcgen05.mma M = 128, N = 160, K = 16, split-bf16 x3, into one
//               of TWO TMEM accumulators, so the mainloop of tile i+1 runs under the epilogue of tile i;
//   warps 2-9   epilogue: a thread owns one logits row (= TMEM lane) and one 80-column half of it. Pass 1 walks the accumulator
//               columns out of TMEM, adds the bias, keeps the fp32 value in a per-row shared-memory scratch and pushes an
//               order-preserving integer key (low 8 bits = column) through a branch-free max/min insertion network - no shuffles,
//               no divergence, 7 integer ops per logit for K <= 4; pass 2 re-reads the scratch for the sum of exponentials and
//               fetches the exact fp32 logits of the K winners; the upper half hands its (keys, logits, max, sum) to the lower
//               half through the scratch row; one record of 16-byte vectors per (row, tile) goes to global memory.
// MEGA = true is the whole modified_beam_search time loop in ONE cooperative launch (cfg4: 250 frames): four more warps per CTA
// run the hypothesis merge + next-operand step (beam_merge.cuh) of the CTA's streams, and the two kinds of work are chained by
// per-frame, per-row-tile counters in global memory (release / acquire) instead of kernel boundaries: a stream's merge of frame t
// starts when the 35 column tiles of its row tile are reduced, a tile of frame t+1 is loaded when the 32 streams of its row tile
// have written their operand rows. Tiles are walked row-major, so the merges of the first wave's row tiles run under the second
// wave's GEMMs and vice versa.
// MEGA = false (beams 3 / 5 / 6 / 7, or no room for a cooperative launch): one launch per frame with programmatic stream
// serialization - barrier set-up, TMEM allocation and the first weight copies overlap the tail of the kernel before;
// griddepcontrol.wait precedes the first read of x and every write.
#include <stdlib.h>

#include <type_traits>

#include "k2b_internal.h"
#include "sm100_ptx.cuh"
#include "beam_merge.cuh"

namespace k2b {

namespace {

using namespace ptx;

constexpr int kJM = 128, kJNc = 160, kJBK = 64, kStages = 2;
constexpr int kATile = kJM * 128;                          // 16 KB: 128 rows x 64 bf16
// kJNc is the width of the packed weight image's tiles; a kernel instantiation works on BN = 160 or 80 columns (an 80-column tile
// is one contiguous half of every 160-row image block): the narrow one doubles the number of tiles when there are too few of them
// to occupy the SMs (cfg3: 4 x 13 wide tiles), halving the GEMM and epilogue latency of the frame chain
template <int BN> struct Tile {
  static constexpr int kWTile = BN * 128;                          // 20 KB / 10 KB
  static constexpr int kStageBytes = 2 * kATile + 2 * kWTile;      // 72 KB / 52 KB
  static constexpr int kTS = BN + 1;                               // row stride of the fp32 scratch rows (conflict-free across rows)
  static constexpr int kTileBytes = kJM * kTS * 4;
};
__device__ __forceinline__ size_t w_image_offset(int col0, int nkb, int kb) {      // byte offset of rows col0.. of k-block kb
  return (((size_t)(col0 / kJNc) * nkb + kb) * kJNc + (size_t)(col0 % kJNc)) * 128;
}
constexpr int kEpiWarps = 8;                             // epilogue warps per CTA (4: one thread per row; 8: two column halves)
constexpr int kAccCols = 256;                             // TMEM column distance of the two accumulators
constexpr int kNoneKey = (int)0x80000000;
constexpr int kInfKey = (int)0x807fffff;                  // keys of -inf logits (invalid columns) are <= this

struct JArgs {
  const uint8_t* a_img;      // [M/128][J/64][hi 16 KB | lo 16 KB]
  const uint8_t* w_hi_img;   // [ntn][J/64][20 KB]
  const uint8_t* w_lo_img;
  const float* bias;         // [V]
  int M, ntm, ntn, nkb, x3, nvalid, topk;
  float* part_rec;           // [M,ntn,kBeamRecWords<KK>]: (max, sum-exp, KK values, KK indices) per (row, tile)
  int* status;
  long long* dbg;
  long long* tl;             // diagnostic timeline of this launch: [sm][8] clock64 stamps (slots 0..3), or null
  // MEGA: the whole search
  int B, V, T, blank, unk, mask3, J;
  int t0, Ttot;              // this launch steps frames t0 .. t0 + T - 1 of a search over Ttot frames (time chunks of a host call)
  BeamState st[2];           // frame t reads st[t & 1], writes st[(t & 1) ^ 1]
  int32_t* bp;
  const int32_t* lens;
  const float* dec_tab;
  const float* enc;          // [B,T,J] projected frames
  long long enc_stride;
  uint8_t* x_img;            // = a_img
  int* done;                 // [T][ntm]: column tiles of (frame, row tile) whose partials are written
  int* ready;                // [T+1][ntm]: streams of (frame, row tile) whose operand rows are written
  int* abort_flag;
  uint32_t epoch0;           // MEGA, != 0: TAGGED records - those of frame t carry epoch0 + t + 1 (beam 1: in word 0; beams 4 / 8: in
                             // the fourth word of every 16-byte vector, beam_merge.cuh) and no `done` counter is kept
  GreedyOut go;              // beam 1: where the merge warps leave tokens / timestamps / counts (/ Hyp)
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded wait for *p >= need (another CTA's release); false on time-out or when some CTA has given up
__device__ __forceinline__ bool wait_count(const int* p, int need, int* abort_flag) {
  if (ld_acquire_gpu(p) >= need) return true;
  const long long t0 = clock64();
  int spins = 0;
#pragma unroll 1
  while (clock64() - t0 < kWaitTimeoutCycles) {
    if (ld_acquire_gpu(p) >= need) return true;
    if (((++spins) & 63) == 0 && ld_acquire_gpu(abort_flag) != 0) return false;
  }
  atomicExch(abort_flag, 1);
  return false;
}

template <int KK>
__device__ __forceinline__ void push_key(int (&a)[KK], int t) {
#pragma unroll
  for (int i = 0; i < KK; ++i) {
    const int hi = max(a[i], t);
    t = min(a[i], t);
    a[i] = hi;
  }
}

template <int KK, bool MEGA, int EW, int BN>
__global__ void __launch_bounds__(64 + EW * 32 + (MEGA ? 128 : 0), 1) joiner_topk_kernel(const JArgs a) {
  constexpr int kWTile = Tile<BN>::kWTile, kStageBytes = Tile<BN>::kStageBytes, kTS = Tile<BN>::kTS;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages], empty[kStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_t[BN];
  __shared__ float mg_v[MEGA ? KK * KK : 1];
  __shared__ int mg_f[MEGA ? KK * KK : 1], mg_ctx[MEGA ? 2 * KK : 1], mg_bad;
  const int nframes = MEGA ? a.T : 1;
  const int spr = MEGA ? kJM / a.topk : 0;               // streams per row tile

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long c_start = (a.dbg != nullptr || a.tl != nullptr) ? clock64() : 0;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int ntiles = a.ntm * a.ntn;
  const int nkb = a.nkb;
  const uint32_t x3 = (uint32_t)a.x3;

  if (tid == 0) {
    mg_bad = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EW); }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t_d = tmem_slot;
  bool ok = true;
  if (!MEGA) griddep_launch_dependents();
  long long* tl = nullptr;
  long long* tlm = (MEGA && a.tl != nullptr && blockIdx.x == 0) ? a.tl : nullptr;   // MEGA: CTA 0 stamps [frame][16] for 40 frames
  if (!MEGA && a.tl != nullptr) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tl = a.tl + (size_t)smid * 8;
    if (tid == 64) tl[0] = c_start;
  }

  if (warp_u == 0) {
    // ---- loader ------------------------------------------------------------------------------------------------------
    if (lane == 0) {
      const uint32_t w_bytes = (uint32_t)(x3 ? 2 * kWTile : kWTile), a_bytes = (uint32_t)(x3 ? 2 * kATile : kATile);
      uint32_t kc = 0;
      bool waited = MEGA;
      for (int t = 0; t < nframes && ok; ++t) {
        for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
          const int tile_m = tile / a.ntn, tile_n = tile - tile_m * a.ntn;      // row-major: a row tile completes early
          bool w_ahead = false;
          if (MEGA && t > 0) {       // the operand rows of this row tile are the merge step's output of frame t - 1
            const int need = min(spr, a.B - tile_m * spr);
            const int* cnt = a.ready + (size_t)t * a.ntm + tile_m;
            if (ld_acquire_gpu(cnt) < need) {
              // not there yet: the weight half of the first k-block does not depend on it - fetch it while waiting
              const int s = kc % kStages;
              const uint32_t ph = (kc / kStages) & 1u;
              if (!mbar_wait(&empty[s], ph ^ 1u)) ok = false;
              uint8_t* st = smem + (size_t)s * kStageBytes;
              const size_t off = w_image_offset(tile_n * BN, nkb, 0);
              mbar_expect_tx(&full[s], w_bytes + a_bytes);
              tma_bulk_g2s(st + 2 * kATile, a.w_hi_img + off, kWTile, &full[s]);
              if (x3) tma_bulk_g2s(st + 2 * kATile + kWTile, a.w_lo_img + off, kWTile, &full[s]);
              w_ahead = true;
              if (!wait_count(cnt, need, a.abort_flag)) { ok = false; break; }
            }
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          if (tlm != nullptr && t < 40) tlm[t * 16 + (tile == blockIdx.x ? 0 : 3)] = clock64();
          for (int kb = 0; kb < nkb; ++kb, ++kc) {
            const int s = kc % kStages;
            const uint32_t ph = (kc / kStages) & 1u;
            uint8_t* st = smem + (size_t)s * kStageBytes;
            if (!(w_ahead && kb == 0)) {
              if (!mbar_wait(&empty[s], ph ^ 1u)) ok = false;
              const size_t off = w_image_offset(tile_n * BN, nkb, kb);
              mbar_expect_tx(&full[s], w_bytes + a_bytes);
              tma_bulk_g2s(st + 2 * kATile, a.w_hi_img + off, kWTile, &full[s]);
              if (x3) tma_bulk_g2s(st + 2 * kATile + kWTile, a.w_lo_img + off, kWTile, &full[s]);
            }
            if (!waited) { griddep_wait(); waited = true; }      // x is the previous kernel's output; the weights are not
            tma_bulk_g2s(st, a.a_img + ((size_t)tile_m * nkb + kb) * (2 * kATile), a_bytes, &full[s]);
          }
        }
      }
    }
  } else if (warp_u == 1) {
    // ---- MMA issuer ----------------------------------------------------------------------------------------------------
    const uint32_t el = elect_one();
    const uint32_t idesc = umma_idesc_bf16_f32(kJM, BN);
    const uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);
    uint32_t kc = 0;
    int it = 0;
    for (int t = 0; t < nframes && ok; ++t)
    for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      if (!mbar_wait(&acc_empty[as], aph ^ 1u)) ok = false;
      tc_fence_after();
      const uint32_t d = t_d + (uint32_t)(as * kAccCols);
      uint32_t acc = 0;
      for (int kb = 0; kb < nkb; ++kb, ++kc) {
        const int s = kc % kStages;
        const uint32_t ph = (kc / kStages) & 1u;
        if (!mbar_wait(&full[s], ph)) ok = false;
        tc_fence_after();
        const uint32_t sb = smem_u32(smem + (size_t)s * kStageBytes);
        const uint32_t ah = (((sb) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t al = (((sb + kATile) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t wh = (((sb + 2 * kATile) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t wl = (((sb + 2 * kATile + kWTile) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dah = ((uint64_t)desc_hi << 32) | (uint64_t)(ah + 2u * k);
          const uint64_t dwh = ((uint64_t)desc_hi << 32) | (uint64_t)(wh + 2u * k);
          umma_ss_e(d, dah, dwh, idesc, acc, el);
          acc = 1;
          if (x3) {
            const uint64_t dal = ((uint64_t)desc_hi << 32) | (uint64_t)(al + 2u * k);
            const uint64_t dwl = ((uint64_t)desc_hi << 32) | (uint64_t)(wl + 2u * k);
            umma_ss_e(d, dah, dwl, idesc, 1, el);
            umma_ss_e(d, dal, dwh, idesc, 1, el);
          }
        }
        umma_commit_e(&empty[s], el);       // frees the stage once these MMAs have read it
      }
      umma_commit_e(&acc_full[as], el);
    }
  } else if (warp_u < 2 + EW) {
    // ---- epilogue: thread = logits row (EW = 4) or one 80-column half of it (EW = 8) ------------------------------------------
    const int lg = warp & 3;                              // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;                     // column half of this warp (always 0 when EW = 4)
    constexpr int kCols = BN * 4 / EW;                    // columns per thread: 160, 80 or 40
    const int cbeg = half * kCols;
    const int row = lg * 32 + lane;
    const int etid = tid - 64;
    float* myrow = reinterpret_cast<float*>(smem + (size_t)kStages * kStageBytes) + (size_t)row * kTS;
    const bool dbg = a.dbg != nullptr && blockIdx.x == 0 && tid == 64;
    long long c_acc = 0, c_epi = 0;
    const int K = a.topk;
    int it = 0;
    if (!MEGA) griddep_wait();                            // partials are read by the previous frame's merge kernel
    if (tl != nullptr && tid == 64) tl[1] = clock64();
    for (int t = 0; t < nframes && ok; ++t)
    for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x, ++it) {
      const int tile_m = tile / a.ntn, tile_n = tile - tile_m * a.ntn;
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const int col0 = tile_n * BN;
      named_bar_sync(1, EW * 32);                         // every row is done with the previous tile's bias (and scratch records)
      for (int c = etid; c < BN; c += EW * 32) bias_t[c] = col0 + c < a.nvalid ? __ldg(a.bias + col0 + c) : -INFINITY;
      named_bar_sync(1, EW * 32);
      if (!mbar_wait(&acc_full[as], aph)) ok = false;
      if (dbg && it == 0) c_acc = clock64();
      if (tl != nullptr && tid == 64 && it == 0) tl[2] = clock64();
      if (tlm != nullptr && tid == 64 && t < 40) tlm[t * 16 + (tile == blockIdx.x ? 1 : 4)] = clock64();
      tc_fence_after();
      const uint32_t trow = t_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * kAccCols);
      int key[KK];
#pragma unroll
      for (int i = 0; i < KK; ++i) key[i] = kNoneKey;
      // pass 1 over a chunk of NC accumulator columns starting at c0: + bias, scratch copy, key network
      auto chunk = [&](const uint32_t* u, int c0, auto nc_tag) {
        constexpr int NC = decltype(nc_tag)::value;
#pragma unroll
        for (int q = 0; q < NC / 4; ++q) {
          const float4 bb = *reinterpret_cast<const float4*>(bias_t + c0 + 4 * q);
          const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            const float v = __uint_as_float(u[j]) + bv[e];
            myrow[c0 + j] = v;
            const int kb = __float_as_int(v);
            const int ok_key = kb ^ ((kb >> 31) & 0x7fffffff);
            push_key<KK>(key, (ok_key & ~255) | (c0 + j));
          }
        }
      };
#pragma unroll 1
      for (int c0 = cbeg; c0 + 32 <= cbeg + kCols; c0 += 32) {
        uint32_t u[32];
        tmem_ld32(trow + (uint32_t)c0, u);
        tmem_ld_wait();
        chunk(u, c0, std::integral_constant<int, 32>{});
      }
      if constexpr (kCols % 32 == 16) {                   // 80 = 2 x 32 + 16
        uint32_t u[16];
        const int c0 = cbeg + kCols - 16;
        tmem_ld16(trow + (uint32_t)c0, u);
        tmem_ld_wait();
        chunk(u, c0, std::integral_constant<int, 16>{});
      } else if constexpr (kCols % 32 == 8) {             // 40 = 32 + 8
        uint32_t u[8];
        const int c0 = cbeg + kCols - 8;
        tmem_ld8(trow + (uint32_t)c0, u);
        tmem_ld_wait();
        chunk(u, c0, std::integral_constant<int, 8>{});
      }
      if (tlm != nullptr && tid == 64 && t < 40 && tile == blockIdx.x) tlm[t * 16 + 10] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);         // the accumulator may be overwritten by tile it + 2
      __syncwarp();
      // pass 2: sum of exponentials against the (key-precision) maximum, then the winners' exact logits
      bool any = key[0] > kInfKey;
      const int mk = key[0] & ~255;
      float mx = any ? __int_as_float(mk ^ ((mk >> 31) & 0x7fffffff)) : -INFINITY;
      const float mneg = any ? -mx * 1.4426950408889634f : 0.f;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      if constexpr (KK > 1)                               // beam 1 (greedy): the log-softmax is a constant shift, nobody reads the sum
#pragma unroll 8
      for (int j = cbeg; j < cbeg + kCols; j += 4) {
        float e0, e1, e2, e3;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(myrow[j], 1.4426950408889634f, mneg)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(myrow[j + 1], 1.4426950408889634f, mneg)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(myrow[j + 2], 1.4426950408889634f, mneg)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fmaf(myrow[j + 3], 1.4426950408889634f, mneg)));
        s0 += e0; s1 += e1; s2 += e2; s3 += e3;
      }
      float sum = any ? (KK > 1 ? (s0 + s1) + (s2 + s3) : 1.f) : 0.f;     // beam 1: any positive constant (see above)
      if (tlm != nullptr && tid == 64 && t < 40 && tile == blockIdx.x) tlm[t * 16 + 11] = clock64();
      float tv[KK];
#pragma unroll
      for (int i = 0; i < KK; ++i) tv[i] = key[i] > kInfKey ? myrow[key[i] & 255] : -INFINITY;
      if constexpr (EW == 8) {
        // the upper column half leaves (keys, exact logits, max, sum) at the start of its own scratch segment; the lower half folds
        // them into its own: key network again, the logits follow their keys by equality (keys are unique)
        int* rec = reinterpret_cast<int*>(myrow + kCols);
        if (half == 1) {
#pragma unroll
          for (int i = 0; i < KK; ++i) { rec[i] = key[i]; rec[KK + i] = __float_as_int(tv[i]); }
          rec[2 * KK] = __float_as_int(mx);
          rec[2 * KK + 1] = __float_as_int(sum);
        }
        named_bar_sync(4 + lg, 64);
        if (half == 0) {
          int ka[KK], kb2[KK];
          float ta[KK], tb[KK];
#pragma unroll
          for (int i = 0; i < KK; ++i) { ka[i] = key[i]; ta[i] = tv[i]; kb2[i] = rec[i]; tb[i] = __int_as_float(rec[KK + i]); }
          const float mb = __int_as_float(rec[2 * KK]), sb = __int_as_float(rec[2 * KK + 1]);
#pragma unroll
          for (int i = 0; i < KK; ++i) push_key<KK>(key, kb2[i]);
#pragma unroll
          for (int i = 0; i < KK; ++i) {
            float v = -INFINITY;
#pragma unroll
            for (int j = 0; j < KK; ++j) { v = key[i] == ka[j] ? ta[j] : v; v = key[i] == kb2[j] ? tb[j] : v; }
            tv[i] = key[i] > kInfKey ? v : -INFINITY;
          }
          const float mm = fmaxf(mx, mb);
          float ea = 0.f, eb = 0.f;
          if (mm > -INFINITY) {
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"((mx - mm) * 1.4426950408889634f));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"((mb - mm) * 1.4426950408889634f));
          }
          sum = (mx > -INFINITY ? sum * ea : 0.f) + (mb > -INFINITY ? sb * eb : 0.f);
          mx = mm;
          any = mm > -INFINITY;
        }
      }
      const int m = tile_m * kJM + row;
      if (m < a.M && half == 0) {
        constexpr int RW = kBeamRecWords<KK>;
        uint32_t w[RW];
#pragma unroll
        for (int i = 0; i < RW; ++i) w[i] = 0u;
        const bool tagged = MEGA && a.epoch0 != 0u;
        const uint32_t tagv = a.epoch0 + (uint32_t)t + 1u;
        w[0] = (tagged && KK == 1) ? tagv : __float_as_uint(mx);       // beam 1: nobody reads max / sum
        w[1] = __float_as_uint(any ? sum : 0.f);
#pragma unroll
        for (int i = 0; i < KK; ++i) {
          const bool has = i < K && key[i] > kInfKey;
          w[2 + i] = __float_as_uint(has ? tv[i] : -INFINITY);
          w[2 + KK + i] = (uint32_t)(has ? col0 + (key[i] & 255) : -1);
        }
        if (KK > 1 && tagged) {            // (three data words, tag) per vector; indices as 16-bit halves
          uint32_t d[3 * (RW / 4)];
#pragma unroll
          for (int j = 0; j < 3 * (RW / 4); ++j) d[j] = 0u;
          d[0] = w[0]; d[1] = w[1];
#pragma unroll
          for (int i = 0; i < KK; ++i) d[2 + i] = w[2 + i];
#pragma unroll
          for (int i = 0; i < KK; i += 2) d[2 + KK + (i >> 1)] = (w[2 + KK + i] & 0xffffu) | (w[2 + KK + i + 1] << 16);
#pragma unroll
          for (int v = 0; v < RW / 4; ++v) { w[4 * v] = d[3 * v]; w[4 * v + 1] = d[3 * v + 1]; w[4 * v + 2] = d[3 * v + 2]; w[4 * v + 3] = tagv; }
        }
        uint4* dst = reinterpret_cast<uint4*>(a.part_rec + ((size_t)m * a.ntn + tile_n) * RW);
#pragma unroll
        for (int v = 0; v < RW / 4; ++v) dst[v] = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
      }
      if (dbg && it == 0) c_epi = clock64();
      if (tlm != nullptr && tid == 64 && t < 40 && tile == blockIdx.x) tlm[t * 16 + 12] = clock64();
      if (MEGA && a.epoch0 == 0u) {                       // publish: one more column tile of (frame, row tile) is reduced
                                                          // (tagged records announce themselves)
        named_bar_sync(3, EW * 32);                       // (CTA barrier, then one gpu-scope release: cumulative over the CTA's stores)
        if (etid == 0) red_release_gpu(a.done + (size_t)t * a.ntm + tile_m, 1);
        if (tlm != nullptr && tid == 64 && t < 40) tlm[t * 16 + (tile == blockIdx.x ? 2 : 5)] = clock64();
      }
    }
    if (tl != nullptr && tid == 64) tl[3] = clock64();
    if (dbg) {
      const long long c_end = clock64();
      atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg + 10), (unsigned long long)(c_acc - c_start));
      atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg + 11), (unsigned long long)(c_epi - c_acc));
      atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg + 12), (unsigned long long)(c_end - c_start));
      atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg + 13), 1ull);
    }
  }
  else if (MEGA) {
    // ---- merge warps: hypothesis merge + next operand of this CTA's streams, frame by frame -------------------------------------
    const int mtid = tid - (64 + EW * 32);
    if constexpr (KK == 1) {
      // greedy (beam 1): a warp steps a stream on its own - the CTA's streams are dealt to its four merge warps
      const int mw = mtid >> 5;
      for (int t = 0; t < nframes && ok; ++t) {
        const float* enc_next = t + 1 < a.T ? a.enc + (size_t)(t + 1) * a.J : nullptr;
        const bool odd = ((a.t0 + t) & 1) != 0;
        BeamState sin, sout;
        sin.ctx = odd ? a.st[1].ctx : a.st[0].ctx; sout.ctx = odd ? a.st[0].ctx : a.st[1].ctx;
        sin.lp = odd ? a.st[1].lp : a.st[0].lp; sout.lp = odd ? a.st[0].lp : a.st[1].lp;
        sin.len = odd ? a.st[1].len : a.st[0].len; sout.len = odd ? a.st[0].len : a.st[1].len;
        sin.hash = odd ? a.st[1].hash : a.st[0].hash; sout.hash = odd ? a.st[0].hash : a.st[1].hash;
        sin.nlive = odd ? a.st[1].nlive : a.st[0].nlive; sout.nlive = odd ? a.st[0].nlive : a.st[1].nlive;
        for (int s = blockIdx.x + mw * gridDim.x; s < a.B && ok; s += 4 * gridDim.x) {
          const int r = s / spr;
          if (a.epoch0 == 0u) {                           // records announced by the row tile's counter
            int good = 1;
            if (lane == 0) good = wait_count(a.done + (size_t)t * a.ntm + r, a.ntn, a.abort_flag) ? 1 : 0;
            good = __shfl_sync(0xffffffffu, good, 0);
            if (!good) { ok = false; break; }
          }
          if (!greedy_merge_warp(lane, s, a.V, a.ntn, a.Ttot, a.t0 + t, a.blank, a.unk, a.mask3, a.part_rec, sin, sout, a.go, a.lens,
                                 a.dec_tab, enc_next, a.enc_stride, a.J, a.x_img, a.epoch0 != 0u ? a.epoch0 + (uint32_t)t + 1u : 0u,
                                 a.abort_flag)) {
            if (lane == 0) atomicExch(a.abort_flag, 1);
            ok = false;
            break;
          }
          if (t + 1 < a.T) {
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncwarp();
            if (lane == 0) red_release_gpu(a.ready + (size_t)(t + 1) * a.ntm + r, 1);
          }
        }
      }
    } else
    for (int t = 0; t < nframes && ok; ++t) {
      const float* enc_next = t + 1 < a.T ? a.enc + (size_t)(t + 1) * a.J : nullptr;
      for (int s = blockIdx.x; s < a.B && ok; s += gridDim.x) {
        const int r = s / spr;
        float4 pe0, pe1;
        beam_merge_prefetch(mtid, s, a.topk, a.J, enc_next, a.enc_stride, &pe0, &pe1);
        if (a.epoch0 == 0u) {                             // counters (more than 64 column tiles, or indices beyond 16 bits)
          int good = 1;
          if (mtid == 0) good = wait_count(a.done + (size_t)t * a.ntm + r, a.ntn, a.abort_flag) ? 1 : 0;
          if (mtid == 0) mg_ctx[0] = good;
          named_bar_sync(2, 128);
          good = mg_ctx[0];
          named_bar_sync(2, 128);                         // mg_ctx is rewritten by the merge step
          if (!good) { ok = false; break; }
        }
        if (tlm != nullptr && mtid == 0 && t < 40) tlm[t * 16 + (s == blockIdx.x ? 6 : 8)] = clock64();
        const bool odd = ((a.t0 + t) & 1) != 0;
        BeamState sin, sout;
        sin.ctx = odd ? a.st[1].ctx : a.st[0].ctx; sout.ctx = odd ? a.st[0].ctx : a.st[1].ctx;
        sin.lp = odd ? a.st[1].lp : a.st[0].lp; sout.lp = odd ? a.st[0].lp : a.st[1].lp;
        sin.len = odd ? a.st[1].len : a.st[0].len; sout.len = odd ? a.st[0].len : a.st[1].len;
        sin.hash = odd ? a.st[1].hash : a.st[0].hash; sout.hash = odd ? a.st[0].hash : a.st[1].hash;
        sin.nlive = odd ? a.st[1].nlive : a.st[0].nlive; sout.nlive = odd ? a.st[0].nlive : a.st[1].nlive;
        beam_merge_stream<KK>(mtid, 2, s, a.topk, a.V, a.ntn, a.Ttot, a.t0 + t, a.blank, a.unk, a.mask3, a.part_rec,
                              sin, sout, a.bp, a.lens, a.dec_tab, enc_next, a.enc_stride, a.J, a.x_img, pe0, pe1,
                              mg_v, mg_f, mg_ctx, (tlm != nullptr && t < 40 && s == blockIdx.x) ? tlm + 640 + t * 8 : nullptr,
                              a.epoch0 != 0u ? a.epoch0 + (uint32_t)t + 1u : 0u, a.abort_flag, &mg_bad);
        if (a.epoch0 != 0u) {
          named_bar_sync(2, 128);
          if (mg_bad != 0) { if (mtid == 0) atomicExch(a.abort_flag, 1); ok = false; break; }
        }
        if (t + 1 < a.T) {                                // publish: one more stream of (frame t + 1, row tile) has its operand rows
          asm volatile("fence.proxy.async;" ::: "memory");       // the rows are read through the async proxy (TMA)
          named_bar_sync(2, 128);
          if (mtid == 0) red_release_gpu(a.ready + (size_t)(t + 1) * a.ntm + r, 1);
        }
        if (tlm != nullptr && mtid == 0 && t < 40) tlm[t * 16 + (s == blockIdx.x ? 7 : 9)] = clock64();
      }
    }
  }
  if (!ok) atomicExch(a.status, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(t_d, 512);
}

template <int KK, bool MEGA, int BN>
int32_t launch_bn(k2b_handle* h, const JArgs& a) {
  static bool attr_set = false;
  const size_t smem = (size_t)kStages * Tile<BN>::kStageBytes + Tile<BN>::kTileBytes;
  if (!attr_set) {
    K2B_CUDA(h, (cudaFuncSetAttribute(joiner_topk_kernel<KK, MEGA, kEpiWarps, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    attr_set = true;
  }
  const int tiles = a.ntm * a.ntn;
  // MEGA: CTAs beyond the tile count still merge their share of the streams
  const int want = MEGA && a.B > tiles ? a.B : tiles;
  // (a caller that projects the next chunk's frames on a side stream meanwhile leaves SMs free: mega_grid_cap)
  const int sms = MEGA && h->mega_grid_cap > 0 && h->mega_grid_cap >= tiles && h->mega_grid_cap < h->sm_count ? h->mega_grid_cap : h->sm_count;
  const int grid = want < sms ? want : sms;
  if (MEGA) {
    // every CTA waits for counters other CTAs publish: all of them must be resident at once (cooperative launch checks that)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + kEpiWarps * 32 + 128); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (h->la_mark_arm) {
      K2B_CUDA(h, cudaEventRecord(h->la_ev_mark, h->stream));
      h->la_mark_valid = true;
    }
    const cudaError_t e = cudaLaunchKernelEx(&cfg, joiner_topk_kernel<KK, MEGA, kEpiWarps, BN>, a);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
      cudaGetLastError();                 // not all CTAs can be resident (SMs taken by another context): per-frame launches instead
      return kMegaUnavailable;
    }
    K2B_CUDA(h, e);
  } else {
    K2B_CUDA(h, (launch_pdl(joiner_topk_kernel<KK, MEGA, kEpiWarps, BN>, dim3(grid), dim3(64 + kEpiWarps * 32), smem, h->stream, a)));
  }
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

template <int KK, bool MEGA>
int32_t launch_as(k2b_handle* h, const JArgs& a, int bn) {
  return bn == 80 ? launch_bn<KK, MEGA, 80>(h, a) : launch_bn<KK, MEGA, kJNc>(h, a);
}

}  // namespace

bool joiner_topk_supported(const k2b_handle* h, int topk) {
  return topk >= 1 && topk <= 8 && h->cfg.joiner_dim % kJBK == 0 && h->wj_ready;
}

bool joiner_topk_usable(const k2b_handle* h, int topk) {
  return topk >= 1 && topk <= 8 && h->cfg.joiner_dim % kJBK == 0;
}

// Tile width for M hypothesis rows: 80 columns when the 160-column tiling leaves more than half of the SMs without a tile
int joiner_topk_width(const k2b_handle* h, int M) {
  const int tiles = ((M + kJM - 1) / kJM) * ((h->cfg.vocab_size + kJNc - 1) / kJNc);
  return 2 * tiles <= h->sm_count ? 80 : kJNc;
}
int joiner_topk_tiles(const k2b_handle* h, int M) {
  const int bn = joiner_topk_width(h, M);
  return (h->cfg.vocab_size + bn - 1) / bn;
}

// x_img: the joiner operand as bf16 hi / lo tile images (joinin_table_tc / decoder_joinin_tc). Partial records per (row, 160-column
// vocabulary tile): beam_partial_words(topk) floats each (beam_merge.cuh). The weight images must exist (ensure_joiner_assets).
int32_t joiner_topk_tc(k2b_handle* h, const uint8_t* x_img, int M, int topk, int kk, float* part_rec) {
  h->ll_clean_ptr = nullptr;             // untagged records: the persistent greedy kernel zeroes the buffer before its next use
  JArgs a = {};
  a.a_img = x_img; a.w_hi_img = h->wj_hi_img; a.w_lo_img = h->wj_lo_img; a.bias = h->out_b;
  const int bn = joiner_topk_width(h, M);
  a.M = M; a.ntm = (M + kJM - 1) / kJM; a.ntn = (h->cfg.vocab_size + bn - 1) / bn; a.nkb = h->cfg.joiner_dim / kJBK;
  a.x3 = h->cfg.precision == K2B_PREC_BF16 ? 0 : 1;
  a.nvalid = h->cfg.vocab_size; a.topk = topk;
  a.part_rec = part_rec;
  a.status = h->dev_status + 1;
  a.dbg = h->cluster_timing;
  a.tl = h->timeline != nullptr ? h->timeline + (size_t)(h->timeline_frame % 64) * 148 * 8 : nullptr;
  return kk == 1 ? launch_as<1, false>(h, a, bn) : kk == 4 ? launch_as<4, false>(h, a, bn) : launch_as<8, false>(h, a, bn);
}


// ints of the per-(frame, row tile) counters + the abort flag of one persistent launch (zeroed before it)
size_t beam_mega_sync_ints(const k2b_handle* h, int B, int T, int K) {
  (void)h;
  const int ntm = (B * K + kJM - 1) / kJM;
  return (size_t)(2 * T + 1) * ntm + 1;
}

// ---- the whole modified_beam_search time loop in one launch (memoised decoder, K in {2, 4, 8}) --------------------------------
bool beam_mega_usable(const k2b_handle* h, int K) {
  const bool off = h->opt_no_mega != 0;                    // k2b_set_option("no_mega"): tests compare the two engines in one process
  return !off && (K == 1 || K == 2 || K == 4 || K == 8) && joiner_topk_usable(h, K) && h->dec_tab != nullptr;
}

int32_t beam_mega_tc(k2b_handle* h, const float* enc, int B, int T, int K, uint8_t* x_img, float* part_rec, const BeamStatePtrs& s0,
                     const BeamStatePtrs& s1, int32_t* bp, const int32_t* lens, int mask3, int kk, const GreedyOutPtrs* go, bool sync_zeroed,
                     int t0, int Ttot, long long enc_stride) {
  const int M = B * K;
  JArgs a = {};
  a.a_img = x_img; a.x_img = x_img; a.w_hi_img = h->wj_hi_img; a.w_lo_img = h->wj_lo_img; a.bias = h->out_b;
  const int bn = joiner_topk_width(h, M);
  a.M = M; a.ntm = (M + kJM - 1) / kJM; a.ntn = (h->cfg.vocab_size + bn - 1) / bn; a.nkb = h->cfg.joiner_dim / kJBK;
  a.x3 = h->cfg.precision == K2B_PREC_BF16 ? 0 : 1;
  a.nvalid = h->cfg.vocab_size; a.topk = K;
  a.part_rec = part_rec;
  a.status = h->dev_status + 1;
  a.B = B; a.V = h->cfg.vocab_size; a.T = T; a.blank = h->cfg.blank_id; a.unk = h->cfg.unk_id; a.mask3 = mask3; a.J = h->cfg.joiner_dim;
  a.st[0] = BeamState{s0.ctx, s0.lp, s0.len, reinterpret_cast<uint64_t*>(s0.hash), s0.nlive};
  a.st[1] = BeamState{s1.ctx, s1.lp, s1.len, reinterpret_cast<uint64_t*>(s1.hash), s1.nlive};
  a.bp = bp; a.lens = lens; a.dec_tab = h->dec_tab; a.enc = enc; a.enc_stride = enc_stride;
  a.t0 = t0; a.Ttot = Ttot;
  const size_t nsync = beam_mega_sync_ints(h, B, T, K);
  K2B_TRY(ensure(h, h->ws_sync, nsync * sizeof(int)));
  if (!sync_zeroed) K2B_CUDA(h, cudaMemsetAsync(h->ws_sync.p, 0, nsync * sizeof(int), h->stream));
  if (kk == 1) {
    if (go == nullptr) return fail(h, K2B_ERR_INVALID, "beam_mega_tc: beam 1 writes its tokens itself and needs the output arrays");
    a.go = GreedyOut{go->tokens, go->ts, go->n, go->hyp, go->cap};
  }
  // Measured (tools/time_mega_tagged.py): greedy, cfg3's shape: 11.95 -> 11.0 us per frame with tagged records; beam 4, cfg4: 20.8 ->
  // 22.8 us when all 128 merge threads poll every vector of their records (the joiner's weight stream already runs at the L2 -> SM
  // limit), 20.8 -> 20.2 us when ONE warp polls one word per record with a pause between rounds (beam_merge_stream).
  if (h->opt_tagged_records != 0 && (kk == 1 || (a.ntn <= 64 && h->cfg.vocab_size < 65535))) {
    // tagged records: whatever another engine (or another layout) left in the partials buffer must not look like a tag
    if (h->ll_clean_ptr != h->ws_part.p || h->ll_clean_bytes != h->ws_part.bytes || h->ll_clean_kk != kk) {
      K2B_CUDA(h, cudaMemsetAsync(h->ws_part.p, 0, h->ws_part.bytes, h->stream));
      h->ll_clean_ptr = h->ws_part.p; h->ll_clean_bytes = h->ws_part.bytes; h->ll_clean_kk = kk;
      h->ll_epoch = 1;
    }
    a.epoch0 = h->ll_epoch;
    h->ll_epoch += (uint32_t)T + 1u;
    if (h->ll_epoch > 0xfff00000u) { h->ll_clean_ptr = nullptr; }       // long before the tags wrap: start again from a zeroed buffer
  } else {
    h->ll_clean_ptr = nullptr;           // records of another layout go into the same buffer
  }
  a.done = static_cast<int*>(h->ws_sync.p);
  a.ready = a.done + (size_t)T * a.ntm;
  a.abort_flag = a.ready + (size_t)(T + 1) * a.ntm;
  a.tl = h->timeline;
  return kk == 1 ? launch_as<1, true>(h, a, bn) : kk == 4 ? launch_as<4, true>(h, a, bn) : launch_as<8, true>(h, a, bn);
}


// kk = compile-time beam bound of the instantiation: 1 (greedy: top-1 records without sums), 4 or 8
int beam_partial_words(int kk) { return kk == 1 ? kBeamRecWords<1> : kk == 4 ? kBeamRecWords<4> : kBeamRecWords<8>; }

}  // namespace k2b
