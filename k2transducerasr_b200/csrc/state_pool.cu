// On-device streaming state: per-stream slots of encoder caches and the per-stream <-> batched re-layout around an encoder call.
//
// replaces: stack_states / unstack_states (ref OnlineProjOfZipformer2.cs:144-362 / :363-489 and the Zipformer, Lstm and Conformer
// siblings), which interleave managed arrays with Array.Copy: for every cache tensor i, with A = its "axisnum" and X = len/A,
//     stacked_i[(x*B + n)*A + a] = item_{n,i}[x*A + a]            (batch on axis 1 of [X, B, A])
// Here a stream's caches live concatenated in one slot of a device pool; stacking B slots is one launch that reads and writes
// every byte once (HBM-bound: 2 * B * Ls * 4 bytes), coalesced along `a`, 16 bytes per thread when the layout allows it.
#include "k2b_internal.h"

namespace k2b {

namespace {

constexpr int kChunk = 1024;        // per-stream elements of one tensor handled by one CTA (for all B streams)

struct TensorDesc {
  int off;      // offset of the tensor inside a slot (floats)
  int len;      // per-stream length
  int axis;     // A
  int first;    // index of its first chunk in the chunk table
};

// UNSTACK = false: pool -> stacked; true: stacked -> pool
template <bool UNSTACK>
__global__ void __launch_bounds__(256) restack_kernel(float* __restrict__ pool, size_t slot_stride, const int32_t* __restrict__ slots,
                                                      int B, const TensorDesc* __restrict__ td, const int2* __restrict__ chunks,
                                                      float* __restrict__ stacked) {
  const int2 ck = chunks[blockIdx.x];              // (tensor, first per-stream element)
  const TensorDesc t = td[ck.x];
  const int e1 = min(ck.y + kChunk, t.len);
  float* sbase = stacked + (size_t)B * t.off;      // tensor i of the stacked set starts at B * off_i
  const bool vec = (t.axis % 4 == 0) && (t.off % 4 == 0) && (slot_stride % 4 == 0);
  if (vec) {
    // 16 bytes per thread, four streams per trip so that four independent loads are in flight per thread
    const int e = ck.y + 4 * (int)threadIdx.x;            // kChunk == 4 * blockDim.x: one vector per thread and stream
    if (e >= e1) return;
    const int x = e / t.axis, a = e - x * t.axis;
    for (int n0 = blockIdx.y * 4; n0 < B; n0 += gridDim.y * 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = n0 + u;
        if (n < B) {
          const float4* src = UNSTACK ? reinterpret_cast<const float4*>(sbase + ((size_t)x * B + n) * t.axis + a)
                                      : reinterpret_cast<const float4*>(pool + (size_t)slots[n] * slot_stride + t.off + e);
          v[u] = *src;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = n0 + u;
        if (n < B) {
          float4* dst = UNSTACK ? reinterpret_cast<float4*>(pool + (size_t)slots[n] * slot_stride + t.off + e)
                                : reinterpret_cast<float4*>(sbase + ((size_t)x * B + n) * t.axis + a);
          *dst = v[u];
        }
      }
    }
    return;
  }
  for (int n = blockIdx.y; n < B; n += gridDim.y) {
    float* item = pool + (size_t)slots[n] * slot_stride + t.off;
    {
      for (int e = ck.y + (int)threadIdx.x; e < e1; e += (int)blockDim.x) {
        const int x = e / t.axis, a = e - x * t.axis;
        float* s = sbase + ((size_t)x * B + n) * t.axis + a;
        if (UNSTACK) item[e] = *s; else *s = item[e];
      }
    }
  }
}

}  // namespace

struct StatePool {
  float* mem = nullptr;
  size_t slot_stride = 0;      // floats
  int max_streams = 0;
  std::vector<int> len, off;
  TensorDesc* d_td = nullptr;
  int2* d_chunks = nullptr;
  int32_t* d_slots = nullptr;
  int slots_cap = 0;
  std::vector<int> axis_cached;   // axis set the device tables were built for
  int nchunks = 0;
};

void state_pool_free(k2b_handle* h) {
  StatePool* p = h->state_pool;
  if (p == nullptr) return;
  if (p->mem) cudaFree(p->mem);
  if (p->d_td) cudaFree(p->d_td);
  if (p->d_chunks) cudaFree(p->d_chunks);
  if (p->d_slots) cudaFree(p->d_slots);
  delete p;
  h->state_pool = nullptr;
}

int32_t state_pool_create(k2b_handle* h, const int32_t* item_len, int n_tensors, int max_streams) {
  state_pool_free(h);
  StatePool* p = new StatePool();
  h->state_pool = p;
  size_t total = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (item_len[i] <= 0) return fail(h, K2B_ERR_INVALID, "k2b_state_pool_create: tensor lengths must be positive");
    p->off.push_back((int)total);
    p->len.push_back(item_len[i]);
    total += (size_t)((item_len[i] + 3) & ~3);          // every tensor starts 16-byte aligned inside a slot
    if (total > 0x7fffffffull) return fail(h, K2B_ERR_INVALID, "k2b_state_pool_create: a stream's state exceeds 2^31 floats");
  }
  p->slot_stride = total;
  p->max_streams = max_streams;
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->mem), sizeof(float) * total * (size_t)max_streams));
  K2B_CUDA(h, cudaMemsetAsync(p->mem, 0, sizeof(float) * total * (size_t)max_streams, h->stream));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->d_td), sizeof(TensorDesc) * (size_t)n_tensors));
  return K2B_OK;
}

static int32_t prepare_tables(k2b_handle* h, StatePool* p, const int32_t* axis_len) {
  const int nt = (int)p->len.size();
  bool same = (int)p->axis_cached.size() == nt;
  for (int i = 0; same && i < nt; ++i) same = p->axis_cached[i] == axis_len[i];
  if (same) return K2B_OK;
  std::vector<TensorDesc> td((size_t)nt);
  std::vector<int2> chunks;
  for (int i = 0; i < nt; ++i) {
    if (axis_len[i] <= 0 || p->len[i] % axis_len[i] != 0)
      return fail(h, K2B_ERR_INVALID, "k2b_stack_states: axis_len must divide the tensor length");
    td[(size_t)i] = TensorDesc{p->off[i], p->len[i], axis_len[i], (int)chunks.size()};
    for (int e = 0; e < p->len[i]; e += kChunk) chunks.push_back(make_int2(i, e));
  }
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));          // earlier launches may still read the old tables
  if (p->d_chunks) cudaFree(p->d_chunks);
  p->d_chunks = nullptr;
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->d_chunks), sizeof(int2) * chunks.size()));
  K2B_CUDA(h, cudaMemcpy(p->d_td, td.data(), sizeof(TensorDesc) * td.size(), cudaMemcpyHostToDevice));
  K2B_CUDA(h, cudaMemcpy(p->d_chunks, chunks.data(), sizeof(int2) * chunks.size(), cudaMemcpyHostToDevice));
  p->nchunks = (int)chunks.size();
  p->axis_cached.assign(axis_len, axis_len + nt);
  return K2B_OK;
}

int32_t state_pool_restack(k2b_handle* h, const int32_t* slots, int B, const int32_t* axis_len, float* stacked_dev, bool unstack) {
  StatePool* p = h->state_pool;
  if (p == nullptr) return fail(h, K2B_ERR_STATE, "k2b_stack_states: no state pool (call k2b_state_pool_create)");
  if (B <= 0) return K2B_OK;
  for (int n = 0; n < B; ++n) {
    if (slots[n] < 0 || slots[n] >= p->max_streams) return fail(h, K2B_ERR_INVALID, "k2b_stack_states: slot out of range");
    if (unstack)
      for (int m = 0; m < n; ++m)
        if (slots[m] == slots[n]) return fail(h, K2B_ERR_INVALID, "k2b_unstack_states: a slot appears twice");
  }
  K2B_TRY(prepare_tables(h, p, axis_len));
  if (p->slots_cap < B) {
    K2B_CUDA(h, cudaStreamSynchronize(h->stream));
    if (p->d_slots) cudaFree(p->d_slots);
    p->d_slots = nullptr;
    K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->d_slots), sizeof(int32_t) * (size_t)B));
    p->slots_cap = B;
  }
  K2B_CUDA(h, cudaMemcpyAsync(p->d_slots, slots, sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
  // grid.x = chunks of the per-stream layout, grid.y splits the B streams so that small pools still fill the GPU
  int gy = (4 * h->sm_count + p->nchunks - 1) / p->nchunks;
  gy = gy < 1 ? 1 : (gy > (B + 3) / 4 ? (B + 3) / 4 : gy);
  dim3 grid((unsigned)p->nchunks, (unsigned)gy);
  if (unstack) restack_kernel<true><<<grid, 256, 0, h->stream>>>(p->mem, p->slot_stride, p->d_slots, B, p->d_td, p->d_chunks, stacked_dev);
  else restack_kernel<false><<<grid, 256, 0, h->stream>>>(p->mem, p->slot_stride, p->d_slots, B, p->d_td, p->d_chunks, stacked_dev);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

int32_t state_pool_io(k2b_handle* h, int slot, float* host, bool put) {
  StatePool* p = h->state_pool;
  if (p == nullptr) return fail(h, K2B_ERR_STATE, "k2b_state_pool: no state pool (call k2b_state_pool_create)");
  if (slot < 0 || slot >= p->max_streams) return fail(h, K2B_ERR_INVALID, "k2b_state_pool: slot out of range");
  float* base = p->mem + (size_t)slot * p->slot_stride;
  size_t hoff = 0;
  for (size_t i = 0; i < p->len.size(); ++i) {            // the host side is the plain concatenation, the slot pads to 16 bytes
    if (put) K2B_CUDA(h, cudaMemcpyAsync(base + p->off[i], host + hoff, sizeof(float) * (size_t)p->len[i], cudaMemcpyHostToDevice, h->stream));
    else K2B_CUDA(h, cudaMemcpyAsync(host + hoff, base + p->off[i], sizeof(float) * (size_t)p->len[i], cudaMemcpyDeviceToHost, h->stream));
    hoff += (size_t)p->len[i];
  }
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  return K2B_OK;
}

size_t state_pool_stacked_floats(const k2b_handle* h, int B) {
  return h->state_pool ? h->state_pool->slot_stride * (size_t)B : 0;
}

}  // namespace k2b
