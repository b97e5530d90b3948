// Hardware self-test of the tcgen05 building blocks the persistent search kernel relies on. It pins, on a real
// B200, the encodings written out in sm100_ptx.cuh: K-major SWIZZLE_128B shared-memory operand layout and
// descriptor, instruction descriptor, TMEM accumulator layout as seen by tcgen05.ld, the TMEM-resident A operand
// (tcgen05.st packed bf16 pairs + tcgen05.mma with A in TMEM), bulk TMA copy + mbarrier transaction count, and a
// cluster round trip through distributed shared memory.
#include <string.h>

#include "k2b_internal.h"
#include "sm100_ptx.cuh"

namespace k2b {

namespace {

using namespace ptx;

constexpr int kM = 128;

// mode 0: D = bf16(A) * bf16(B)^T                          (SS)
// mode 1: D = Ah*Bh + Ah*Bl + Al*Bh, all operands in smem  (SS x3)
// mode 2: same, with Al resident in TMEM                   (SS, SS, TS)
// mode 3: D = bf16(A) * bf16(B)^T with A resident in TMEM  (TS)
// `Apk` (optional): A_hi pre-packed by the host in the exact swizzled smem image, fetched with one bulk TMA copy.
__global__ void __launch_bounds__(128)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, const uint8_t* __restrict__ Apk, int N, int K,
                     int mode, float* __restrict__ D, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = K / 64;
  uint8_t* a_hi = smem;                                  // nkb tiles of 128 x 128 B
  uint8_t* a_lo = a_hi + (size_t)nkb * 16384;
  uint8_t* b_hi = a_lo + (size_t)nkb * 16384;            // nkb tiles of N x 128 B
  uint8_t* b_lo = b_hi + (size_t)nkb * N * 128;
  __shared__ uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_tma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t t_d = tbase;            // 32 columns of accumulator
  const uint32_t t_a = tbase + 64;       // up to 256 columns of packed A
  const uint32_t lane_base = (uint32_t)(32 * warp) << 16;

  if (Apk != nullptr && tid == 0) {
    mbar_expect_tx(&bar_tma, (uint32_t)(nkb * 16384));
    tma_bulk_g2s(a_hi, Apk, (uint32_t)(nkb * 16384), &bar_tma);
  }
  // A: thread = row
  {
    const float* row = A + (size_t)tid * K;
    for (int k0 = 0; k0 < K; k0 += 64) {
      uint32_t pk_hi[32], pk_lo[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x0 = row[k0 + 2 * j], x1 = row[k0 + 2 * j + 1];
        const float h0 = bf16_round(x0), h1 = bf16_round(x1);
        pk_hi[j] = pack_bf16x2(h0, h1);
        pk_lo[j] = pack_bf16x2(x0 - h0, x1 - h1);
      }
      uint8_t* th = a_hi + (size_t)(k0 / 64) * 16384;
      uint8_t* tl = a_lo + (size_t)(k0 / 64) * 16384;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (Apk == nullptr) *reinterpret_cast<uint32_t*>(th + sw128_offset(tid, 2 * j)) = pk_hi[j];
        *reinterpret_cast<uint32_t*>(tl + sw128_offset(tid, 2 * j)) = pk_lo[j];
      }
      if (mode == 2) tmem_st32(t_a + lane_base + (uint32_t)(k0 / 2), pk_lo);
      if (mode == 3) tmem_st32(t_a + lane_base + (uint32_t)(k0 / 2), pk_hi);
    }
    if (mode >= 2) tmem_st_wait();
  }
  if (tid < N) {
    const float* row = B + (size_t)tid * K;
    for (int k = 0; k < K; k += 2) {
      const float x0 = row[k], x1 = row[k + 1];
      const float h0 = bf16_round(x0), h1 = bf16_round(x1);
      const size_t tile = (size_t)(k / 64) * N * 128;
      *reinterpret_cast<uint32_t*>(b_hi + tile + sw128_offset(tid, k & 63)) = pack_bf16x2(h0, h1);
      *reinterpret_cast<uint32_t*>(b_lo + tile + sw128_offset(tid, k & 63)) = pack_bf16x2(x0 - h0, x1 - h1);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  bool ok = true;
  if (tid == 0) {
    if (Apk != nullptr) ok = mbar_wait(&bar_tma, 0);
    const uint32_t idesc = umma_idesc_bf16_f32(kM, N);
    uint32_t acc = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      for (int k = 0; k < 4; ++k) {
        const uint64_t dah = umma_desc_k_sw128(smem_u32(a_hi + (size_t)kb * 16384) + k * 32);
        const uint64_t dal = umma_desc_k_sw128(smem_u32(a_lo + (size_t)kb * 16384) + k * 32);
        const uint64_t dbh = umma_desc_k_sw128(smem_u32(b_hi + (size_t)kb * N * 128) + k * 32);
        const uint64_t dbl = umma_desc_k_sw128(smem_u32(b_lo + (size_t)kb * N * 128) + k * 32);
        const uint32_t ta = t_a + (uint32_t)((kb * 4 + k) * 8);
        if (mode == 3) {
          umma_ts(t_d, ta, dbh, idesc, acc); acc = 1;
        } else {
          umma_ss(t_d, dah, dbh, idesc, acc); acc = 1;
          if (mode >= 1) {
            umma_ss(t_d, dah, dbl, idesc, 1);
            if (mode == 1) umma_ss(t_d, dal, dbh, idesc, 1);
            else umma_ts(t_d, ta, dbh, idesc, 1);
          }
        }
      }
    }
    umma_commit(&bar_mma);
  }
  if (!mbar_wait(&bar_mma, 0)) ok = false;
  tc_fence_after();
  if (!ok) atomicExch(status, 1);
  uint32_t v[32];
  tmem_ld32(t_d + lane_base, v);
  tmem_ld_wait();
  for (int n = 0; n < N && n < 32; ++n) D[(size_t)tid * N + n] = __uint_as_float(v[n]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// CTA-pair MMA (cta_group::2): D[256,N] = bf16(A[256,K]) * bf16(B[N,K])^T on a cluster of two CTAs. CTA r stages rows
// 128r.. of A (shared memory, or TMEM with ts != 0) and rows (N/2)r.. of B; the leader issues, both read their own half of D.
__global__ void __launch_bounds__(128)
umma2_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, int K, int ts, float* __restrict__ D,
                      int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = K / 64, NH = N / 2;
  uint8_t* a_s = smem;                                   // nkb tiles of 128 x 128 B
  uint8_t* b_s = a_s + (size_t)nkb * 16384;              // nkb tiles of NH x 128 B
  __shared__ uint64_t bar_mma, bar_peer;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_peer, 1);
    mbar_fence_init();
  }
  cluster_sync();
  if (warp == 0) tmem_alloc2(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t t_d = tbase, t_a = tbase + 256;
  const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
  {
    const float* row = A + ((size_t)rank * 128 + tid) * K;
    for (int k0 = 0; k0 < K; k0 += 64) {
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) pk[j] = pack_bf16x2(bf16_round(row[k0 + 2 * j]), bf16_round(row[k0 + 2 * j + 1]));
      uint8_t* th = a_s + (size_t)(k0 / 64) * 16384;
#pragma unroll
      for (int j = 0; j < 32; ++j) *reinterpret_cast<uint32_t*>(th + sw128_offset(tid, 2 * j)) = pk[j];
      if (ts) tmem_st32(t_a + lane_base + (uint32_t)(k0 / 2), pk);
    }
    if (ts) tmem_st_wait();
  }
  if (tid < NH) {
    const float* row = B + ((size_t)rank * NH + tid) * K;
    for (int k = 0; k < K; k += 2)
      *reinterpret_cast<uint32_t*>(b_s + (size_t)(k / 64) * NH * 128 + sw128_offset(tid, k & 63)) =
          pack_bf16x2(bf16_round(row[k]), bf16_round(row[k + 1]));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  // the peer tells the leader that its operands are in place (remote arrive, cluster-scope release)
  if (rank == 1 && tid == 0) mbar_arrive_remote(dsmem_map(smem_u32(&bar_peer), 0));
  bool ok = true;
  if (rank == 0 && warp == 0) {
    if (!mbar_wait_cluster(&bar_peer, 0)) ok = false;
    tc_fence_after();
    const uint32_t el = elect_one();
    const uint32_t idesc = umma_idesc_bf16_f32(256, N);
    uint32_t acc = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(a_s + (size_t)kb * 16384) + k * 32);
        const uint64_t db = umma_desc_k_sw128(smem_u32(b_s + (size_t)kb * NH * 128) + k * 32);
        if (ts) umma2_ts_e(t_d, t_a + (uint32_t)((kb * 4 + k) * 8), db, idesc, acc, el);
        else umma2_ss_e(t_d, da, db, idesc, acc, el);
        acc = 1;
      }
    }
    umma2_commit_e(&bar_mma, (uint16_t)0x3, el);
  }
  if (!mbar_wait(&bar_mma, 0)) ok = false;
  tc_fence_after();
  if (!ok) atomicExch(status, 1);
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(t_d + lane_base + (uint32_t)c0, v);
    tmem_ld_wait();
    for (int n = 0; n < 32 && c0 + n < N; ++n) D[((size_t)rank * 128 + tid) * N + c0 + n] = __uint_as_float(v[n]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) tmem_dealloc2(tbase, 512);
}

// MMA timing microbenchmark: `reps` back-to-back k-loops (nkb*4 MMAs each) of one flavour, cycles from first issue to
// commit completion. flavour 0: SS N=32, 1: SS N=64, 2: TS N=32, 3: TS N=64, 4: SS N=32 with M=64, 5: SS N=128, 6: TS N=128,
// 7/8/9: SS N=32 rotating over 2/4/8 independent TMEM accumulators, 10: SS N=64 over 4 accumulators
__global__ void __launch_bounds__(128) umma_bench_kernel(int flavour, int nkb, int reps, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* a_s = smem;                       // nkb x 16 KB
  uint8_t* b_s = smem + (size_t)nkb * 16384; // nkb x 16 KB (up to 128 rows)
  for (int i = tid; i < nkb * 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 7);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  const int nacc = flavour == 7 ? 2 : (flavour == 8 ? 4 : (flavour == 9 ? 8 : (flavour == 10 ? 4 : 1)));
  const int N = (flavour == 1 || flavour == 3 || flavour == 10) ? 64 : ((flavour == 5 || flavour == 6) ? 128 : 32);
  const int M = flavour == 4 ? 64 : 128;
  const bool ts = flavour == 2 || flavour == 3 || flavour == 6;
  const uint32_t idesc = umma_idesc_bf16_f32(M, N);
  const uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);
  const uint32_t a0 = ((smem_u32(a_s) & 0x3FFFFu) >> 4) | (1u << 16), b0 = ((smem_u32(b_s) & 0x3FFFFu) >> 4) | (1u << 16);
  const int wu = __shfl_sync(0xffffffffu, warp, 0);      // warp-uniform warp index
  if (flavour >= 11 && wu == 3) {
    // warp-uniform issue loop (flavour 11: SS N=32, 12: TS N=32, 13: SS N=64): descriptors live in uniform registers
    const uint32_t el = elect_one();
    const int N2 = (flavour == 13 || flavour == 14 || flavour == 15) ? 64 : 32;
    const uint32_t idesc2 = umma_idesc_bf16_f32(128, N2);
    const uint32_t idesc64u = umma_idesc_bf16_f32(128, 64), idesc32u = umma_idesc_bf16_f32(128, 32);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a0 + (uint32_t)(kb * 1024 + k * 2));
          const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b0 + (uint32_t)(kb * 1024 + k * 2));
          if (flavour == 12) umma_ts_e(tb, tb + 384 + (uint32_t)((kb * 4 + k) * 8), db, idesc2, 1, el);
          else if (flavour == 14) umma_ss_e(tb + (uint32_t)(((kb * 4 + k) & 1) * 64), da, db, idesc2, 1, el);      // 2 accumulators
          else if (flavour == 15) umma_ss_e(tb + (uint32_t)(((kb * 4 + k) & 3) * 64), da, db, idesc2, 1, el);      // 4 accumulators
          else if (flavour == 16) {            // the kernel's pair, same accumulator: SS N=64 then TS N=32
            umma_ss_e(tb, da, db, idesc64u, 1, el);
            umma_ts_e(tb, tb + 384 + (uint32_t)((kb * 4 + k) * 8), db, idesc32u, 1, el);
          } else if (flavour == 17) {          // the pair on separate accumulators
            umma_ss_e(tb, da, db, idesc64u, 1, el);
            umma_ts_e(tb + 64, tb + 384 + (uint32_t)((kb * 4 + k) * 8), db, idesc32u, 1, el);
          } else if (flavour == 18) {          // the pair, separate accumulators, K split over two more
            const uint32_t o = (uint32_t)((k & 1) * 96);
            umma_ss_e(tb + o, da, db, idesc64u, 1, el);
            umma_ts_e(tb + o + 64, tb + 384 + (uint32_t)((kb * 4 + k) * 8), db, idesc32u, 1, el);
          }
          else umma_ss_e(tb, da, db, idesc2, 1, el);
        }
      }
    }
    long long t1 = clock64();
    umma_commit_e(&bar, el);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (el) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  if (flavour < 11 && tid == 0) {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a0 + (uint32_t)(kb * 1024 + k * 2));
          const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b0 + (uint32_t)(kb * 1024 + k * 2));
          const uint32_t td = tb + (uint32_t)(((kb * 4 + k) % nacc) * 64);   // rotate over independent accumulators
          if (ts) umma_ts(td, tb + 384 + (uint32_t)((kb * 4 + k) * 8), db, idesc, 1);
          else umma_ss(td, da, db, idesc, 1);
        }
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// latency of dependent warp collectives: out[0] = cycles per __reduce_max_sync, out[1] = per 5-level shfl_xor max,
// out[2] = per __ballot_sync, out[3] = per __reduce_max_sync when 16 warps run the same chain concurrently (out[4]: shfl)
__global__ void __launch_bounds__(512) collective_bench_kernel(long long* __restrict__ out, int* __restrict__ sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = lane * 7 + 3;
  const int reps = 256;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < reps; ++i) x = __reduce_max_sync(0xffffffffu, x ^ i) + lane;
  long long t1 = clock64();
  for (int i = 0; i < reps; ++i) {
    int y = x ^ i;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) y = max(y, __shfl_xor_sync(0xffffffffu, y, o));
    x = y + lane;
  }
  long long t2 = clock64();
  for (int i = 0; i < reps; ++i) x = (int)__ballot_sync(0xffffffffu, (x ^ i) & 1) + lane;
  long long t3 = clock64();
  if (threadIdx.x == 0) { out[3] = (t1 - t0) / reps; out[4] = (t2 - t1) / reps; out[5] = (t3 - t2) / reps; }
  __syncthreads();
  if (warp == 0) {       // alone on the SM
    long long a0 = clock64();
    for (int i = 0; i < reps; ++i) x = __reduce_max_sync(0xffffffffu, x ^ i) + lane;
    long long a1 = clock64();
    for (int i = 0; i < reps; ++i) {
      int y = x ^ i;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) y = max(y, __shfl_xor_sync(0xffffffffu, y, o));
      x = y + lane;
    }
    long long a2 = clock64();
    for (int i = 0; i < reps; ++i) x = (int)__ballot_sync(0xffffffffu, (x ^ i) & 1) + lane;
    long long a3 = clock64();
    if (lane == 0) { out[0] = (a1 - a0) / reps; out[1] = (a2 - a1) / reps; out[2] = (a3 - a2) / reps; }
  }
  sink[threadIdx.x] = x;
}

// every CTA writes (rank+1)*1000 + its own tid into slot [rank] of every CTA of the cluster, then checks its slots
__global__ void __launch_bounds__(64) cluster_selftest_kernel(int* __restrict__ out) {
  __shared__ uint32_t slots[16][64];
  const uint32_t rank = cluster_ctarank(), n = cluster_nctarank();
  cluster_sync();                      // all CTAs of the cluster have started: remote smem is valid
  for (uint32_t r = 0; r < n; ++r)
    dsmem_st_u32(dsmem_map(smem_u32(&slots[rank][threadIdx.x]), r), (rank + 1) * 1000 + threadIdx.x);
  cluster_sync();
  int bad = 0;
  for (uint32_t r = 0; r < n; ++r) bad += slots[r][threadIdx.x] != (r + 1) * 1000 + threadIdx.x;
  if (bad) atomicAdd(out, bad);
  if (threadIdx.x == 0) atomicAdd(out + 1, 1);
  cluster_sync();                      // nobody exits while a peer may still write into it
}

// every CTA of the cluster pushes `words16` 16-byte chunks to each peer with st.shared::cluster.v4 (512 threads)
__global__ void __launch_bounds__(512) dsmem_bw_kernel(int words16, long long* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t buf[];
  const uint32_t rank = cluster_ctarank(), n = cluster_nctarank();
  cluster_sync();
  const long long t0 = clock64();
  for (uint32_t p = 1; p < n; ++p) {
    const uint32_t dst = (rank + p) % n;
    const uint32_t base = dsmem_map(smem_u32(buf), dst);
    for (int i = threadIdx.x; i < words16; i += blockDim.x) dsmem_st_v4(base + 16u * i, i, rank, p, 7u);
  }
  cluster_sync();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

}  // namespace

}  // namespace k2b

using namespace k2b;

extern "C" {

// Diagnostic entry point (HOST pointers). A [128,K], B [N,K] fp32; D [128,N]. K multiple of 64, <= 256; N multiple
// of 16, <= 32. mode: see umma_selftest_kernel. use_tma != 0: A_hi arrives pre-swizzled through one bulk TMA copy.
K2B_API int32_t k2b_selftest_umma(k2b_handle* h, const float* A, const float* B, int32_t N, int32_t K, int32_t mode,
                                  int32_t use_tma, float* D) {
  if (h == nullptr) return K2B_ERR_INVALID;
  if (K % 64 || K > 256 || K <= 0 || N % 16 || N > 32 || N <= 0 || mode < 0 || mode > 3)
    return fail(h, K2B_ERR_INVALID, "k2b_selftest_umma: bad shape");
  K2B_CUDA(h, cudaSetDevice(h->cfg.device));
  float *dA, *dB, *dD;
  int* dS;
  uint8_t* dP = nullptr;
  K2B_CUDA(h, cudaMalloc(&dA, sizeof(float) * 128 * K));
  K2B_CUDA(h, cudaMalloc(&dB, sizeof(float) * N * K));
  K2B_CUDA(h, cudaMalloc(&dD, sizeof(float) * 128 * N));
  K2B_CUDA(h, cudaMalloc(&dS, sizeof(int)));
  K2B_CUDA(h, cudaMemcpy(dA, A, sizeof(float) * 128 * K, cudaMemcpyHostToDevice));
  K2B_CUDA(h, cudaMemcpy(dB, B, sizeof(float) * N * K, cudaMemcpyHostToDevice));
  K2B_CUDA(h, cudaMemset(dS, 0, sizeof(int)));
  K2B_CUDA(h, cudaMemset(dD, 0, sizeof(float) * 128 * N));
  const int nkb = K / 64;
  if (use_tma) {
    // host-side pre-pack of A_hi into the swizzled image (same function the kernel uses: row*128 + ((k/8)^(row&7))*16)
    std::vector<uint8_t> img((size_t)nkb * 16384);
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < K; ++k) {
        const float x = A[(size_t)m * K + k];
        uint32_t u;
        memcpy(&u, &x, 4);
        const uint32_t r = ((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
        const uint16_t b = (uint16_t)r;
        const int kk = k & 63;
        const size_t off = (size_t)(k / 64) * 16384 + (size_t)m * 128 + (size_t)(((kk >> 3) ^ (m & 7)) * 16) + (kk & 7) * 2;
        memcpy(&img[off], &b, 2);
      }
    K2B_CUDA(h, cudaMalloc(&dP, img.size()));
    K2B_CUDA(h, cudaMemcpy(dP, img.data(), img.size(), cudaMemcpyHostToDevice));
  }
  const size_t smem = (size_t)nkb * (2 * 16384 + 2 * (size_t)N * 128);
  K2B_CUDA(h, cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, h->stream>>>(dA, dB, dP, N, K, mode, dD, dS);
  K2B_LAUNCH_CHECK(h);
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  int st = 0;
  K2B_CUDA(h, cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
  K2B_CUDA(h, cudaMemcpy(D, dD, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  if (dP) cudaFree(dP);
  if (st != 0) return fail(h, K2B_ERR_STATE, "k2b_selftest_umma: an mbarrier wait timed out");
  return K2B_OK;
}

// cycles[0] = issue loop, cycles[1] = until the commit barrier flips, for reps*nkb*4 MMAs of the given flavour.
// D[256,N] through a CTA pair (cta_group::2). N: multiple of 32, <= 256; K: multiple of 64, <= 256. ts != 0: A from TMEM.
K2B_API int32_t k2b_selftest_umma2(k2b_handle* h, const float* A, const float* B, int32_t N, int32_t K, int32_t ts, float* D) {
  if (h == nullptr || A == nullptr || B == nullptr || D == nullptr || N < 32 || N > 256 || N % 32 || K < 64 || K > 256 || K % 64)
    return K2B_ERR_INVALID;
  K2B_CUDA(h, cudaSetDevice(h->cfg.device));
  float *dA, *dB, *dD; int* st;
  K2B_CUDA(h, cudaMalloc(&dA, sizeof(float) * 256 * K));
  K2B_CUDA(h, cudaMalloc(&dB, sizeof(float) * N * K));
  K2B_CUDA(h, cudaMalloc(&dD, sizeof(float) * 256 * N));
  K2B_CUDA(h, cudaMalloc(&st, sizeof(int)));
  K2B_CUDA(h, cudaMemset(st, 0, sizeof(int)));
  K2B_CUDA(h, cudaMemcpy(dA, A, sizeof(float) * 256 * K, cudaMemcpyHostToDevice));
  K2B_CUDA(h, cudaMemcpy(dB, B, sizeof(float) * N * K, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(K / 64) * (16384 + (size_t)(N / 2) * 128);
  K2B_CUDA(h, cudaFuncSetAttribute(umma2_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  K2B_CUDA(h, cudaLaunchKernelEx(&cfg, umma2_selftest_kernel, (const float*)dA, (const float*)dB, (int)N, (int)K, (int)ts, dD, st));
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  int hs = 0;
  K2B_CUDA(h, cudaMemcpy(&hs, st, sizeof(int), cudaMemcpyDeviceToHost));
  K2B_CUDA(h, cudaMemcpy(D, dD, sizeof(float) * 256 * N, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(st);
  return hs ? fail(h, K2B_ERR_STATE, "cta_group::2 self-test: an mbarrier wait timed out") : K2B_OK;
}

K2B_API int32_t k2b_selftest_umma_bench(k2b_handle* h, int32_t flavour, int32_t nkb, int32_t reps, int64_t* cycles2) {
  if (h == nullptr || nkb < 1 || nkb > 6 || reps < 1) return K2B_ERR_INVALID;
  K2B_CUDA(h, cudaSetDevice(h->cfg.device));
  long long* d;
  K2B_CUDA(h, cudaMalloc(&d, 2 * sizeof(long long)));
  const size_t smem = (size_t)nkb * 32768;
  K2B_CUDA(h, cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_bench_kernel<<<1, 128, smem, h->stream>>>(flavour, nkb, reps, d);
  K2B_LAUNCH_CHECK(h);
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  long long r[2];
  K2B_CUDA(h, cudaMemcpy(r, d, sizeof(r), cudaMemcpyDeviceToHost));
  cudaFree(d);
  cycles2[0] = r[0]; cycles2[1] = r[1];
  return K2B_OK;
}

K2B_API int32_t k2b_selftest_collectives(k2b_handle* h, int64_t* out6) {
  if (h == nullptr) return K2B_ERR_INVALID;
  K2B_CUDA(h, cudaSetDevice(h->cfg.device));
  long long* d; int* sink;
  K2B_CUDA(h, cudaMalloc(&d, 6 * sizeof(long long)));
  K2B_CUDA(h, cudaMalloc(&sink, 512 * sizeof(int)));
  collective_bench_kernel<<<1, 512, 0, h->stream>>>(d, sink);
  K2B_LAUNCH_CHECK(h);
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  long long r[6];
  K2B_CUDA(h, cudaMemcpy(r, d, sizeof(r), cudaMemcpyDeviceToHost));
  cudaFree(d); cudaFree(sink);
  for (int i = 0; i < 6; ++i) out6[i] = r[i];
  return K2B_OK;
}

// cycles for every CTA of `nclusters` clusters of `csize` CTAs to push bytes_per_peer to each peer (incl. two cluster barriers)
K2B_API int32_t k2b_selftest_dsmem_bw(k2b_handle* h, int32_t csize, int32_t nclusters, int32_t bytes_per_peer, int64_t* cycles) {
  if (h == nullptr || csize < 2 || csize > 8 || bytes_per_peer % 16 || bytes_per_peer > 128 * 1024) return K2B_ERR_INVALID;
  K2B_CUDA(h, cudaSetDevice(h->cfg.device));
  long long* d;
  K2B_CUDA(h, cudaMalloc(&d, sizeof(long long)));
  K2B_CUDA(h, cudaFuncSetAttribute(dsmem_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes_per_peer));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize * nclusters); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = bytes_per_peer; cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  K2B_CUDA(h, cudaLaunchKernelEx(&cfg, dsmem_bw_kernel, bytes_per_peer / 16, d));
  h->launches++;
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  long long r;
  K2B_CUDA(h, cudaMemcpy(&r, d, sizeof(r), cudaMemcpyDeviceToHost));
  cudaFree(d);
  *cycles = r;
  return K2B_OK;
}

// Launches `nclusters` clusters of `csize` CTAs; returns the number of mismatching DSMEM slots in *bad.
K2B_API int32_t k2b_selftest_cluster(k2b_handle* h, int32_t csize, int32_t nclusters, int32_t* bad, int32_t* ctas_done) {
  if (h == nullptr || csize < 1 || csize > 16 || nclusters < 1) return K2B_ERR_INVALID;
  K2B_CUDA(h, cudaSetDevice(h->cfg.device));
  int* d;
  K2B_CUDA(h, cudaMalloc(&d, 2 * sizeof(int)));
  K2B_CUDA(h, cudaMemset(d, 0, 2 * sizeof(int)));
  if (csize > 8) K2B_CUDA(h, cudaFuncSetAttribute(cluster_selftest_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize * nclusters);
  cfg.blockDim = dim3(64);
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  K2B_CUDA(h, cudaLaunchKernelEx(&cfg, cluster_selftest_kernel, d));
  h->launches++;
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  int r[2];
  K2B_CUDA(h, cudaMemcpy(r, d, sizeof(r), cudaMemcpyDeviceToHost));
  cudaFree(d);
  if (bad) *bad = r[0];
  if (ctas_done) *ctas_done = r[1];
  return K2B_OK;
}

}  // extern "C"
