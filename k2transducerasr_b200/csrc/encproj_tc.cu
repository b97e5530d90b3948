// encoder_proj on the tensor cores:  C[M,N] = f(A[M,K] * W[N,K]^T + b)   (M = B*T frames, K = E, N = J)
//
// One CTA per 128 x 256 output tile, warp-specialised, 2-stage mbarrier pipeline over 64-wide k-blocks:
//   warp 0      weight loader: bulk TMA copies (cp.async.bulk, SASS UBLKCP) of the pre-split bf16 hi / lo weight
//               tiles, stored at load time in the exact K-major SWIZZLE_128B shared-memory image;
//   warp 1      MMA issuer (warp-uniform loop, one elected lane): tcgen05.mma M=128 N=256 K=16, accumulator in TMEM;
//               split-bf16 x3 (Ah*Wh + Ah*Wl + Al*Wh) for fp32-grade results, or bf16 single pass;
//   warps 2-5   A producers: coalesced fp32 loads of the raw encoder frames, hi/lo bf16 split in registers, swizzled
//               shared-memory stores (the frames are fp32 in HBM, so they are converted on the way in, once);
//               afterwards the same four warps are the epilogue: TMEM -> registers -> +bias -> optional exp(2x)
//               (the form the cluster search kernel's tanh prologue consumes) -> global.
// Algorithmic HBM traffic: M*K*4 (frames in) + M*N*4 (projected frames out); the weights stay in L2.
#include <stdlib.h>

#include "k2b_internal.h"
#include "sm100_ptx.cuh"

namespace k2b {

namespace {

using namespace ptx;

constexpr int kEM = 128, kEN = 256, kBKc = 64, kStages = 2;
// The per-frame joiner uses 160-column vocabulary tiles: 35 x 8 = 280 tiles at V = 5537 run as two waves of small tiles on 148
// SMs, where 256-column tiles (22 x 8 = 176) run as two waves of large ones.
constexpr int kJN = 160;
constexpr int kATile = kEM * 128;        // 16 KB: 128 rows x 64 bf16
constexpr int kWTile = kEN * 128;        // 32 KB
constexpr int kStageBytes = 2 * kATile + 2 * kWTile;   // 96 KB
constexpr int kEThreads = 320;          // loader + MMA issuer + 8 producer warps (the first four of them are also the epilogue)
constexpr int kProducers = 8;

struct EncArgs {
  const float* A;            // [M,K] fp32
  const uint8_t* w_hi_img;   // [N/256][K/64][32 KB]
  const uint8_t* w_lo_img;
  const float* bias;         // [N]
  float* C;                  // [M,N]
  int M, N, K, x3, exp2x;
  int* status;
  // epi 0: store C (optionally exp2x). epi 2 / 3: joiner epilogues - no logits are written, only per (row, 256-column vocab
  // tile) partials: max / sum-exp / top-k (2, modified_beam_search) or the argmax fold with ties and NaN -> larger index (3).
  int epi, nvalid, topk;
  int rows_per_stream, out_T, out_t0;     // epi 0 with rows_per_stream > 0: row m = (b, tt) of a time chunk -> output row b*out_T + out_t0 + tt
  int in_T;                               // > 0 (with rows_per_stream > 0): the INPUT is the whole [B,in_T,K] array too, row (b, out_t0 + tt)
  float* part_m; float* part_s; float* part_tv; int32_t* part_ti;     // epi 2: [M,nt], [M,nt], [M,nt,topk] x2
  float* part_val; int32_t* part_idx; int32_t* part_nan;              // epi 3: [M,nt] each
  // pro 1: stateless decoder as the A producer - A(m,k) = relu(tab0[row(y0(m))][k] + tab1[row(y1(m))][k]) (embedding gather +
  // folded grouped conv1d + ReLU, ref OfflineProjOfTransducer.cs:116 [EXT]); epi 4: C = tanh(acc + bias + enc[m / rows_per_stream])
  // = decoder_proj fused with the joiner prologue (x = tanh(enc + dec))
  int pro, V, neg_wrap;
  const int32_t* ctx; const float* tab0; const float* tab1;
  const float* enc; long long enc_stride;
  // epi 4 can also leave its result as the next GEMM's A operand: bf16 hi / lo tiles in the K-major SWIZZLE_128B shared-memory
  // image, [M/128][N/64][hi 16 KB | lo 16 KB] (x_img != null). pro 2 takes such an image as A: the loader warp fetches the
  // tiles with bulk TMA copies and there is no conversion pass at all - the 22 vocabulary-tile CTAs of a joiner launch used to
  // convert the same fp32 rows 22 times.
  uint8_t* x_img;
  const uint8_t* a_img;
  long long* dbg;            // diagnostic (null in production): CTA 0 adds its phase cycle counts to dbg[10..14]
};

__device__ __forceinline__ int table_row_e(int y, int V, int neg_wrap) {
  if (y >= 0) return y < V ? y : V;      // out-of-range ids read the zero row
  if (neg_wrap) { const int r = y + V; return r >= 0 ? r : V; }
  return V;
}

// tanh(x) = 1 - 2 / (1 + e^(2x)); |x| is clamped where tanh is already 1 to the last bit, so e^(2x) stays finite
__device__ __forceinline__ float tanh_fast(float x) {
  const float y = __expf(2.f * fminf(fmaxf(x, -15.f), 15.f));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y + 1.f));
  return fmaf(-2.f, r, 1.f);
}

__device__ __forceinline__ bool better_e(float v, int i, float ev, int ei) { return v > ev || (v == ev && i > ei); }

// Rare path of the top-k epilogue, kept out of line so that the 256-element scan stays small enough for the instruction
// cache: insert (v, idx) into the sorted list (value desc, index desc) and return the new K-th entry as threshold.
__device__ __noinline__ void topk_insert(float* tv, int* ti, int K, float v, int idx, float* thr_v, int* thr_i) {
  for (int i = 0; i < K; ++i) {
    if (better_e(v, idx, tv[i], ti[i])) {
      const float fv = tv[i]; const int fi = ti[i];
      tv[i] = v; ti[i] = idx; v = fv; idx = fi;
    }
  }
  *thr_v = tv[K - 1];
  *thr_i = ti[K - 1];
}

// EW = warps that run the epilogue: the 8 producer warps, or 16 (the joiner with a TMA-fed A operand has nothing to produce, so
// eight more warps share its reducing epilogue - the longest serial part of a tile).
template <int EW, int BN>
__global__ void __launch_bounds__((2 + EW) * 32, 1) encproj_tc_kernel(const EncArgs a) {
  constexpr int kEN = BN;                                  // tile width of this instantiation (shadows the default)
  constexpr int kWTile = kEN * 128;
  constexpr int kStageBytes = 2 * kATile + 2 * kWTile;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_w[kStages], full_a[kStages], empty[kStages], acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_t[kEN];          // this tile's bias (-inf beyond the valid columns): broadcast reads in the epilogue

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long c_start = a.dbg != nullptr ? clock64() : 0;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int ntn = a.N / kEN;
  const int tile_m = blockIdx.x / ntn, tile_n = blockIdx.x - tile_m * ntn;
  const int nkb = a.K / kBKc;
  const uint32_t x3 = (uint32_t)a.x3;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&full_a[s], kProducers);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  for (int c = tid; c < kEN; c += (2 + EW) * 32) {
    const int col = tile_n * kEN + c;
    bias_t[c] = col < a.nvalid ? __ldg(a.bias + col) : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t_d = tmem_slot;
  bool ok = true;

  if (warp_u == 0) {
    // ---- weight loader -----------------------------------------------------------------------------------
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        if (!mbar_wait(&empty[s], ph ^ 1u)) ok = false;
        uint8_t* st = smem + (size_t)s * kStageBytes;
        const size_t off = ((size_t)tile_n * nkb + kb) * kWTile;
        const uint32_t a_bytes = a.pro == 2 ? (uint32_t)(x3 ? 2 * kATile : kATile) : 0u;
        mbar_expect_tx(&full_w[s], (uint32_t)(x3 ? 2 * kWTile : kWTile) + a_bytes);
        tma_bulk_g2s(st + 2 * kATile, a.w_hi_img + off, kWTile, &full_w[s]);
        if (x3) tma_bulk_g2s(st + 2 * kATile + kWTile, a.w_lo_img + off, kWTile, &full_w[s]);
        if (a.pro == 2) {          // A hi (and lo, adjacent) of this row tile and k-block
          const uint8_t* src = a.a_img + ((size_t)tile_m * nkb + kb) * (2 * kATile);
          tma_bulk_g2s(st, src, a_bytes, &full_w[s]);
        }
      }
    }
  } else if (warp_u == 1) {
    // ---- MMA issuer ------------------------------------------------------------------------------------------
    const uint32_t el = elect_one();
    const uint32_t idesc = umma_idesc_bf16_f32(kEM, kEN);
    const uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);
    uint32_t acc = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
      if (!mbar_wait(&full_w[s], ph)) ok = false;
      if (a.pro != 2 && !mbar_wait(&full_a[s], ph)) ok = false;
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
        umma_ss_e(t_d, dah, dwh, idesc, acc, el);
        acc = 1;
        if (x3) {
          const uint64_t dal = ((uint64_t)desc_hi << 32) | (uint64_t)(al + 2u * k);
          const uint64_t dwl = ((uint64_t)desc_hi << 32) | (uint64_t)(wl + 2u * k);
          umma_ss_e(t_d, dah, dwl, idesc, 1, el);
          umma_ss_e(t_d, dal, dwh, idesc, 1, el);
        }
      }
      umma_commit_e(&empty[s], el);       // frees the stage once these MMAs have read it
    }
    umma_commit_e(&acc_full, el);
  } else {
    // ---- A producers (warps 2..5), then epilogue ------------------------------------------------------------
    const int pw = warp - 2;              // 0..7: rows pw*16 .. pw*16+15 of the tile
    const int half = lane >> 4, c4 = lane & 15;     // two rows per warp instruction, 16 float4 per 64-wide row
    constexpr int kRowsPerWarp = kEM / kProducers, kIters = kRowsPerWarp / 2;
    int r0[kIters], r1[kIters];          // pro 1: table rows of this thread's hypothesis rows
    long long arow[kIters];              // pro 0: input row of tile row i (a time chunk may be read straight out of the whole array)
#pragma unroll
    for (int i = 0; i < kIters; ++i) {
      const int m = tile_m * kEM + pw * kRowsPerWarp + 2 * i + half;
      arow[i] = (a.in_T > 0 && a.rows_per_stream > 0)
                    ? (long long)(m / a.rows_per_stream) * a.in_T + a.out_t0 + (m % a.rows_per_stream) : (long long)m;
    }
    if (a.pro == 1) {
#pragma unroll
      for (int i = 0; i < kIters; ++i) {
        const int m = tile_m * kEM + pw * kRowsPerWarp + 2 * i + half;
        r0[i] = r1[i] = a.V;
        if (m < a.M) {
          r0[i] = table_row_e(a.ctx[2 * m], a.V, a.neg_wrap);
          r1[i] = table_row_e(a.ctx[2 * m + 1], a.V, a.neg_wrap);
        }
      }
    }
    for (int kb = 0; kb < ((a.pro == 2 || pw >= kProducers) ? 0 : nkb); ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
      if (!mbar_wait(&empty[s], ph ^ 1u)) ok = false;
      uint8_t* a_hi = smem + (size_t)s * kStageBytes;
      uint8_t* a_lo = a_hi + kATile;
      float4 v[kIters];
      if (a.pro == 1) {
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(a.tab0 + (size_t)r0[i] * a.K + (size_t)kb * kBKc) + c4);
          const float4 y = __ldg(reinterpret_cast<const float4*>(a.tab1 + (size_t)r1[i] * a.K + (size_t)kb * kBKc) + c4);
          v[i] = make_float4(fmaxf(x.x + y.x, 0.f), fmaxf(x.y + y.y, 0.f), fmaxf(x.z + y.z, 0.f), fmaxf(x.w + y.w, 0.f));
        }
      } else {
#pragma unroll
      for (int i = 0; i < kIters; ++i) {
        const int r = pw * kRowsPerWarp + 2 * i + half;
        const int m = tile_m * kEM + r;
        v[i] = (m < a.M) ? __ldg(reinterpret_cast<const float4*>(a.A + (size_t)arow[i] * a.K + (size_t)kb * kBKc) + c4)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      }
#pragma unroll
      for (int i = 0; i < kIters; ++i) {
        const int r = pw * kRowsPerWarp + 2 * i + half;
        const float h0 = bf16_round(v[i].x), h1 = bf16_round(v[i].y), h2 = bf16_round(v[i].z), h3 = bf16_round(v[i].w);
        const uint32_t off = sw128_offset(r, 4 * c4);
        *reinterpret_cast<uint2*>(a_hi + off) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
        if (x3)
          *reinterpret_cast<uint2*>(a_lo + off) =
              make_uint2(pack_bf16x2(v[i].x - h0, v[i].y - h1), pack_bf16x2(v[i].z - h2, v[i].w - h3));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_a[s]);
    }
    // epilogue (producer warps 0..3 = CTA warps 2..5): a warp owns TMEM lanes 32*(warp%4) .. +31 = those tile rows
    if (a.epi == 2) {
      // ---- softmax / top-k partials of this 128 x 256 logits tile (modified_beam_search), all eight warps -----------
      // (1) accumulator (+bias) -> shared memory, row-major with a 257-word stride (the pipeline stages are dead now);
      // (2) one warp per row, lanes across the 256 columns: selection keys are the order-preserving integer image of the
      //     logit with its low 8 bits replaced by the column, so they are unique and one REDUX per round finds value and
      //     position at once (values closer than 2^-15 relative resolve to the larger vocabulary index; the exact fp32
      //     logit is re-read for the record); sum-exp is a fixed-point REDUX (order-independent).
      const bool dbg = a.dbg != nullptr && blockIdx.x == 0 && tid == 64;
      if (!mbar_wait(&acc_full, 0)) ok = false;
      const long long c_acc = dbg ? clock64() : 0;
      tc_fence_after();
      float* tile = reinterpret_cast<float*>(smem);
      constexpr int kTS = kEN + 1;
      {
        const int lg = warp & 3, cq = pw >> 2;            // EW / 4 column chunks of kEN * 4 / EW columns
        const int row = lg * 32 + lane;
        const uint32_t trow = t_d + ((uint32_t)(lg * 32) << 16);
        constexpr int kChunk = kEN * 4 / EW;               // columns of this warp
        if constexpr (kChunk % 32 == 0) {
          for (int c0 = cq * kChunk; c0 < (cq + 1) * kChunk; c0 += 32) {
            uint32_t u[32];
            tmem_ld32(trow + (uint32_t)c0, u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[row * kTS + c0 + j] = __uint_as_float(u[j]) + bias_t[c0 + j];
          }
        } else {
          static_assert(kChunk % 8 == 0, "tile width / epilogue warps");
          for (int c0 = cq * kChunk; c0 < (cq + 1) * kChunk; c0 += 8) {
            uint32_t u[8];
            tmem_ld8(trow + (uint32_t)c0, u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) tile[row * kTS + c0 + j] = __uint_as_float(u[j]) + bias_t[c0 + j];
          }
        }
      }
      tc_fence_before();
      named_bar_sync(1, EW * 32);
      const long long c_tile = dbg ? clock64() : 0;
      const int K = a.topk;
      const int col0 = tile_n * kEN;
      const int nval = min(kEN, a.nvalid - col0);
      constexpr int kNone = (int)0x80000000;
      constexpr int kRI = 4;             // rows in flight per warp: independent REDUX chains hide each other's latency
      for (int rr = 0; rr < kEM / EW; rr += kRI) {
        constexpr int VPL = kEN / 32;        // logits per lane and row
        float v[kRI][VPL];
        int pk[kRI][VPL];
#pragma unroll
        for (int r = 0; r < kRI; ++r) {
          const int row = pw * (kEM / EW) + rr + r;
#pragma unroll
          for (int j = 0; j < VPL; ++j) {
            const int pos = lane + 32 * j;
            v[r][j] = tile[row * kTS + pos];
            const int kb = __float_as_int(v[r][j]);
            const int key = kb ^ ((kb >> 31) & 0x7fffffff);
            pk[r][j] = pos < nval ? ((key & ~255) | pos) : kNone;
          }
        }
        int keep[kRI];
#pragma unroll
        for (int r = 0; r < kRI; ++r) keep[r] = kNone;
        float mx[kRI], sum[kRI];
        for (int q = 0; q < K; ++q) {
#pragma unroll
          for (int r = 0; r < kRI; ++r) {
            int hk = pk[r][0];
#pragma unroll
            for (int j = 1; j < VPL; ++j) hk = max(hk, pk[r][j]);
            const int wk = __reduce_max_sync(0xffffffffu, hk);
#pragma unroll
            for (int j = 0; j < VPL; ++j) pk[r][j] = (pk[r][j] == wk) ? kNone : pk[r][j];
            keep[r] = (lane == q) ? wk : keep[r];
            if (q == 0) {
              const int mk = wk & ~255;
              mx[r] = (wk == kNone) ? -INFINITY : __int_as_float(mk ^ ((mk >> 31) & 0x7fffffff));
              const float mneg = (wk == kNone) ? 0.f : -mx[r] * 1.4426950408889634f;
              float ls = 0.f;
#pragma unroll
              for (int j = 0; j < VPL; ++j) {
                float ex;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(v[r][j], 1.4426950408889634f, mneg)));
                ls += (lane + 32 * j < nval) ? ex : 0.f;
              }
              const unsigned tot = __reduce_add_sync(0xffffffffu, __float2uint_rn(ls * 8388608.f));
              sum[r] = (wk == kNone) ? 0.f : (float)tot * (1.f / 8388608.f);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < kRI; ++r) {
          const int row = pw * (kEM / EW) + rr + r;
          const int m = tile_m * kEM + row;
          if (m < a.M) {
            const size_t po = (size_t)m * ntn + tile_n;
            if (lane == 0) { a.part_m[po] = mx[r]; a.part_s[po] = sum[r]; }
            if (lane < K) {
              const bool has = keep[r] != kNone;
              const int pos = keep[r] & 255;
              a.part_tv[po * K + lane] = has ? tile[row * kTS + pos] : -INFINITY;
              a.part_ti[po * K + lane] = has ? col0 + pos : -1;
            }
          }
        }
        __syncwarp();
      }
      if (dbg) {
        const long long c_end = clock64();
        atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg + 10), (unsigned long long)(c_acc - c_start));
        atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg + 11), (unsigned long long)(c_tile - c_acc));
        atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg + 12), (unsigned long long)(c_end - c_tile));
        atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg + 13), 1ull);
      }
    } else if (kEN == 256 && (a.epi == 0 || a.epi == 4) && pw < kProducers) {
      // ---- store epilogues, all eight warps: accumulator -> shared memory (row stride 260 words: 16-byte vector accesses both
      //      ways without bank conflicts), then one warp per row with the lanes across the 256 columns, so bias / encoder frame
      //      loads and the output stores are coalesced 512-byte rows
      if (!mbar_wait(&acc_full, 0)) ok = false;
      tc_fence_after();
      float* tile = reinterpret_cast<float*>(smem);
      constexpr int kTS4 = kEN + 4;
      {
        const int lg = warp & 3, chalf = pw >> 2;
        const int row = lg * 32 + lane;
        const uint32_t trow = t_d + ((uint32_t)(lg * 32) << 16);
        for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
          uint32_t u[32];
          tmem_ld32(trow + (uint32_t)c0, u);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(tile + row * kTS4 + c0 + 4 * q) = make_uint4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
        }
      }
      tc_fence_before();
      named_bar_sync(1, kProducers * 32);
      for (int rr = 0; rr < kEM / kProducers; ++rr) {
        const int row = pw * (kEM / kProducers) + rr;
        const int m = tile_m * kEM + row;
        if (m >= a.M) break;
        size_t orow = (size_t)m;
        if (a.epi == 0 && a.rows_per_stream > 0) { const int b = m / a.rows_per_stream; orow = (size_t)b * a.out_T + a.out_t0 + (m - b * a.rows_per_stream); }
        float* crow = a.C != nullptr ? a.C + orow * a.N + (size_t)tile_n * kEN : nullptr;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int col = hf * 128 + 4 * lane;
          float4 o = *reinterpret_cast<const float4*>(tile + row * kTS4 + col);
          const float4 bb = *reinterpret_cast<const float4*>(bias_t + col);
          o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          if (a.epi == 4) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(a.enc + (size_t)(m / a.rows_per_stream) * a.enc_stride +
                                                                    (size_t)tile_n * kEN + col));
            o.x = tanh_fast(o.x + e.x); o.y = tanh_fast(o.y + e.y); o.z = tanh_fast(o.z + e.z); o.w = tanh_fast(o.w + e.w);
          } else if (a.exp2x) {
            o.x = expf(2.f * fminf(fmaxf(o.x, -21.f), 21.f));
            o.y = expf(2.f * fminf(fmaxf(o.y, -21.f), 21.f));
            o.z = expf(2.f * fminf(fmaxf(o.z, -21.f), 21.f));
            o.w = expf(2.f * fminf(fmaxf(o.w, -21.f), 21.f));
          }
          if (a.C != nullptr) *reinterpret_cast<float4*>(crow + col) = o;
          if (a.epi == 4 && a.x_img != nullptr) {
            const int n = tile_n * kEN + col;                        // 4 consecutive K positions of the next GEMM
            uint8_t* timg = a.x_img + ((size_t)tile_m * (a.N / kBKc) + (n >> 6)) * (2 * kATile) + sw128_offset(row, n & 63);
            const float h0 = bf16_round(o.x), h1 = bf16_round(o.y), h2 = bf16_round(o.z), h3 = bf16_round(o.w);
            *reinterpret_cast<uint2*>(timg) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
            *reinterpret_cast<uint2*>(timg + kATile) = make_uint2(pack_bf16x2(o.x - h0, o.y - h1), pack_bf16x2(o.z - h2, o.w - h3));
          }
        }
      }
    } else if (pw < 4) {
    if (!mbar_wait(&acc_full, 0)) ok = false;
    tc_fence_after();
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const int m = tile_m * kEM + row;
    const uint32_t trow = t_d + ((uint32_t)(lg * 32) << 16);
    {
      // the thread owns one hypothesis row and walks its 256 logits of this vocab tile straight out of TMEM
      const int col0 = tile_n * kEN;
      const size_t po = (size_t)m * ntn + tile_n;
      if (a.epi == 3) {
        float bv = 0.f;
        int bi = -1, bnan = 0;
        for (int c0 = 0; c0 < kEN; c0 += 32) {
          uint32_t u[32];
          tmem_ld32(trow + (uint32_t)c0, u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + c0 + j;
            if (col < a.nvalid) {
              const float v = __uint_as_float(u[j]) + bias_t[c0 + j];
              const int vn = (v != v) ? 1 : 0;
              // sequential fold  token = logits[token] > logits[k] ? token : k  (ref OfflineRecognizer.cs:153)
              if (bi < 0 || vn || !(bv > v)) { bv = v; bi = col; if (vn) bnan = 1; }
            }
          }
        }
        if (m < a.M) { a.part_val[po] = bv; a.part_idx[po] = bi; a.part_nan[po] = bnan; }
      }
    }
    }
  }
  if (!ok) atomicExch(a.status, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(t_d, 256);
}

__global__ void pack_enc_w_kernel(const float* __restrict__ w, int N, int K, int BN, uint8_t* __restrict__ hi, uint8_t* __restrict__ lo) {
  const int n = blockIdx.x;                 // weight row
  const int tn = n / BN, r = n % BN, nkb = K / kBKc;
  for (int k = 2 * threadIdx.x; k < K; k += 2 * blockDim.x) {
    const float x0 = w[(size_t)n * K + k], x1 = w[(size_t)n * K + k + 1];
    const float h0 = bf16_round(x0), h1 = bf16_round(x1);
    const size_t off = ((size_t)tn * nkb + (k >> 6)) * ((size_t)BN * 128) + sw128_offset(r, k & 63);
    *reinterpret_cast<uint32_t*>(hi + off) = pack_bf16x2(h0, h1);
    *reinterpret_cast<uint32_t*>(lo + off) = pack_bf16x2(x0 - h0, x1 - h1);
  }
}

}  // namespace

bool encproj_tc_supported(const k2b_handle* h) {
  const k2b_config& c = h->cfg;
  return c.encoder_dim > 0 && c.encoder_dim % 64 == 0 && c.joiner_dim % kEN == 0 && h->enc_w != nullptr;
}

int32_t ensure_encproj_assets(k2b_handle* h) {
  if (h->enc_ready) return K2B_OK;
  const int N = h->cfg.joiner_dim, K = h->cfg.encoder_dim;
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->we_hi_img), (size_t)N * K * 2));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->we_lo_img), (size_t)N * K * 2));
  pack_enc_w_kernel<<<N, 128, 0, h->stream>>>(h->enc_w, N, K, kEN, h->we_hi_img, h->we_lo_img);
  K2B_LAUNCH_CHECK(h);
  h->enc_ready = true;
  return K2B_OK;
}

template <int EW, int BN>
static int32_t launch_tc_as(k2b_handle* h, const EncArgs& a) {
  const int tiles = ((a.M + kEM - 1) / kEM) * (a.N / BN);
  const size_t smem = (size_t)kStages * (2 * kATile + 2 * BN * 128);
  K2B_CUDA(h, cudaFuncSetAttribute(encproj_tc_kernel<EW, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  encproj_tc_kernel<EW, BN><<<tiles, (2 + EW) * 32, smem, h->stream>>>(a);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

// joiner launches (epi 2 / 3) use kJN-column tiles - their weight images are packed that way -, everything else kEN
static int32_t launch_tc(k2b_handle* h, EncArgs& a) {
  a.x3 = h->cfg.precision == K2B_PREC_BF16 ? 0 : 1;
  a.status = h->dev_status + 1;
  if (a.epi == 2 && a.pro == 2) return launch_tc_as<16, kJN>(h, a);
  if (a.epi == 2 || a.epi == 3) return launch_tc_as<8, kJN>(h, a);
  return launch_tc_as<8, kEN>(h, a);
}

// raw [n,E] -> out [n,J] = f(raw * We^T + be) on tcgen05; f = identity or exp(2*clamp(., +-21))
int32_t encoder_proj_tc(k2b_handle* h, const float* raw, int n, float* out, bool exp2x, int rows_per_stream, int out_T, int out_t0,
                        int in_T) {
  K2B_TRY(ensure_encproj_assets(h));
  EncArgs a = {};
  a.rows_per_stream = rows_per_stream; a.out_T = out_T; a.out_t0 = out_t0; a.in_T = in_T;
  a.A = raw; a.w_hi_img = h->we_hi_img; a.w_lo_img = h->we_lo_img; a.bias = h->enc_b; a.C = out;
  a.M = n; a.N = h->cfg.joiner_dim; a.K = h->cfg.encoder_dim;
  a.exp2x = exp2x ? 1 : 0;
  a.epi = 0; a.nvalid = a.N;
  return launch_tc(h, a);
}

// ---- joiner operand from the memoised decoder table: x[m,:] = tanh(enc[m / rows_per_stream] + decoder(ctx[m])) written straight as
//      the bf16 hi / lo tile images the joiner GEMM's loader warp fetches (one block per hypothesis row, one float4 per thread) ----
namespace {
__global__ void joinin_table_kernel(const float* __restrict__ dec_tab, const int32_t* __restrict__ ctx, int M, int V, int J,
                                    const float* __restrict__ enc, long long enc_stride, int rows_per_stream,
                                    uint8_t* __restrict__ x_img) {
  const int m = blockIdx.x, k = 4 * threadIdx.x;
  griddep_launch_dependents();
  if (k >= J) return;
  const float4 e = __ldg(reinterpret_cast<const float4*>(enc + (size_t)(m / rows_per_stream) * enc_stride + k));
  griddep_wait();           // the contexts are the merge kernel's output; the frames are not
  const float4 d = __ldg(reinterpret_cast<const float4*>(dec_tab + ((size_t)(ctx[2 * m] + 1) * V + ctx[2 * m + 1]) * J + k));
  float x[4];
  const float ev[4] = {e.x, e.y, e.z, e.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {           // tanh(e + d) = 1 - 2 / (1 + exp(2e) * exp(2d)); the table holds exp(2d)
    const float y = fmaf(expf(2.f * fminf(fmaxf(ev[i], -21.f), 21.f)), dv[i], 1.f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
    x[i] = fmaf(-2.f, r, 1.f);
  }
  uint8_t* timg = x_img + ((size_t)(m / kEM) * (J / kBKc) + (k >> 6)) * (2 * kATile) + sw128_offset(m % kEM, k & 63);
  const float h0 = bf16_round(x[0]), h1 = bf16_round(x[1]), h2 = bf16_round(x[2]), h3 = bf16_round(x[3]);
  *reinterpret_cast<uint2*>(timg) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
  *reinterpret_cast<uint2*>(timg + kATile) = make_uint2(pack_bf16x2(x[0] - h0, x[1] - h1), pack_bf16x2(x[2] - h2, x[3] - h3));
}
}  // namespace

int32_t joinin_table_tc(k2b_handle* h, const int32_t* ctx, int M, const float* enc, long long enc_stride, int rows_per_stream,
                        uint8_t* x_img) {
  const int J = h->cfg.joiner_dim;
  K2B_CUDA(h, launch_pdl(joinin_table_kernel, dim3(M), dim3((J / 4 + 31) / 32 * 32), 0, h->stream, (const float*)h->dec_tab, ctx, M,
                          h->cfg.vocab_size, J, enc, enc_stride, rows_per_stream, x_img));
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

// ---- stateless decoder + joiner prologue on the tensor cores: x[m,:] = tanh(enc[m / rows_per_stream] + decoder(ctx[m])) -----
bool decoder_tc_supported(const k2b_handle* h) {
  const k2b_config& c = h->cfg;
  return c.decoder_dim % 64 == 0 && c.joiner_dim % kEN == 0 && h->dec_w != nullptr && h->tab0 != nullptr;
}

int32_t decoder_joinin_tc(k2b_handle* h, const int32_t* ctx, int M, const float* enc, long long enc_stride, int rows_per_stream,
                          float* x, uint8_t* x_img) {
  if (!h->wd_ready) {
    const int N = h->cfg.joiner_dim, K = h->cfg.decoder_dim;
    K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->wd_hi_img), (size_t)N * K * 2));
    K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->wd_lo_img), (size_t)N * K * 2));
    pack_enc_w_kernel<<<N, 128, 0, h->stream>>>(h->dec_w, N, K, kEN, h->wd_hi_img, h->wd_lo_img);
    K2B_LAUNCH_CHECK(h);
    h->wd_ready = true;
  }
  EncArgs a = {};
  a.pro = 1; a.ctx = ctx; a.tab0 = h->tab0; a.tab1 = h->tab1; a.V = h->cfg.vocab_size;
  a.neg_wrap = h->cfg.neg_id_mode == K2B_NEGID_WRAP ? 1 : 0;
  a.enc = enc; a.enc_stride = enc_stride; a.rows_per_stream = rows_per_stream;
  a.w_hi_img = h->wd_hi_img; a.w_lo_img = h->wd_lo_img; a.bias = h->dec_b; a.C = x; a.x_img = x_img;
  a.M = M; a.N = h->cfg.joiner_dim; a.K = h->cfg.decoder_dim;
  a.epi = 4; a.nvalid = a.N;
  return launch_tc(h, a);
}

// ---- per-frame tensor-core joiner (any vocabulary): logits tile = x * out_w^T + out_b, reduced in the epilogue ----------
bool joiner_tc_supported(const k2b_handle* h) { return h->cfg.joiner_dim % 64 == 0 && h->out_w != nullptr; }

int joiner_tc_tiles(const k2b_handle* h) { return (h->cfg.vocab_size + kJN - 1) / kJN; }

int32_t ensure_joiner_assets(k2b_handle* h) {
  if (h->wj_ready) return K2B_OK;
  const int V = h->cfg.vocab_size, K = h->cfg.joiner_dim, Np = joiner_tc_tiles(h) * kJN;
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->wj_hi_img), (size_t)Np * K * 2));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->wj_lo_img), (size_t)Np * K * 2));
  K2B_CUDA(h, cudaMemsetAsync(h->wj_hi_img, 0, (size_t)Np * K * 2, h->stream));
  K2B_CUDA(h, cudaMemsetAsync(h->wj_lo_img, 0, (size_t)Np * K * 2, h->stream));
  pack_enc_w_kernel<<<V, 128, 0, h->stream>>>(h->out_w, V, K, kJN, h->wj_hi_img, h->wj_lo_img);
  K2B_LAUNCH_CHECK(h);
  h->wj_ready = true;
  return K2B_OK;
}

// x [M,J] fp32 (already tanh(enc+dec)). topk > 0: softmax/top-k partials; topk == 0: argmax partials. nt = joiner_tc_tiles().
size_t joiner_tc_image_bytes(const k2b_handle* h, int M) {
  return (size_t)((M + kEM - 1) / kEM) * (size_t)(h->cfg.joiner_dim / kBKc) * (2 * kATile);
}

int32_t joiner_tc_partials(k2b_handle* h, const float* x, const uint8_t* x_img, int M, int topk, float* part_m, float* part_s,
                           float* part_tv, int32_t* part_ti, float* part_val, int32_t* part_idx, int32_t* part_nan) {
  K2B_TRY(ensure_joiner_assets(h));
  EncArgs a = {};
  a.A = x; a.w_hi_img = h->wj_hi_img; a.w_lo_img = h->wj_lo_img; a.bias = h->out_b; a.C = nullptr;
  if (x_img != nullptr) { a.pro = 2; a.a_img = x_img; }
  a.M = M; a.N = joiner_tc_tiles(h) * kJN; a.K = h->cfg.joiner_dim;
  a.exp2x = 0;
  a.epi = topk > 0 ? 2 : 3; a.nvalid = h->cfg.vocab_size; a.topk = topk;
  a.part_m = part_m; a.part_s = part_s; a.part_tv = part_tv; a.part_ti = part_ti;
  a.part_val = part_val; a.part_idx = part_idx; a.part_nan = part_nan;
  a.dbg = h->cluster_timing;
  return launch_tc(h, a);
}

}  // namespace k2b
