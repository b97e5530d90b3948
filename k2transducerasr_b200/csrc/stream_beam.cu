// Streaming modified_beam_search: the hypotheses of an OnlineStream carried from chunk to chunk on the device.
//
// The reference's online recognizer accepts `decodingMethod` / `maxActivePaths` (ref OnlineRecognizer.cs:18-19) but only ever
// decodes greedily (ref :46-57); the state it carries between GetResults calls is OnlineStream.Hyp / Tokens / Timestamps
// (ref OnlineStream.cs:44-55, OnlineRecognizer.cs:206-213). Beam search needs more than the last two tokens: per stream the K live
// hypotheses (context, log-prob, length, sequence hash) and the back-pointer history of every frame decoded so far, from which
// the best hypothesis' whole token / timestamp sequence is re-read after each chunk (the best hypothesis may change retroactively).
// That state lives here, in device slots keyed like the encoder-cache pool (state_pool.cu): the search engines run on compact
// per-call arrays, `gather` fills them from the slots of the streams of this call, `scatter` writes the state back and appends the
// chunk's back-pointer rows to each stream's history at its own frame offset, `backtrace` walks the history.
#include <vector>

#include "k2b_internal.h"

namespace k2b {

struct BeamPool {
  int K = 0, max_streams = 0, max_frames = 0;
  int32_t* ctx = nullptr;              // [S][K][2]
  float* lp = nullptr;                 // [S][K]
  int32_t* len = nullptr;              // [S][K]
  unsigned long long* hash = nullptr;  // [S][K]
  int32_t* nlive = nullptr;            // [S]
  int32_t* cst = nullptr;              // [S][K] context-graph (hot word) state
  int32_t* nframes = nullptr;          // [S] frames decoded so far
  int32_t* hist = nullptr;             // [S][max_frames][K] back-pointer rows: (parent slot << 28) | (token + 1)
  int32_t* slots = nullptr;            // [max_streams] the slots of the call in flight
  std::vector<int> nframes_host;       // host mirror (capacity check without a device read)
  std::vector<int> cur_slots;
};

namespace {
constexpr unsigned long long kSeedHash = 0x9E3779B97F4A7C15ull;

__global__ void pool_reset_kernel(int K, int slot, int c0, int c1, int blank, int32_t* __restrict__ ctx, float* __restrict__ lp,
                                  int32_t* __restrict__ len, unsigned long long* __restrict__ hash, int32_t* __restrict__ nlive,
                                  int32_t* __restrict__ nframes, int32_t* __restrict__ cst) {
  const int k = threadIdx.x;
  if (k >= K) return;
  const size_t o = (size_t)slot * K + k;
  cst[o] = 0;
  ctx[2 * o] = k == 0 ? c0 : -1;
  ctx[2 * o + 1] = k == 0 ? c1 : blank;
  lp[o] = k == 0 ? 0.f : -INFINITY;
  len[o] = 2;
  hash[o] = kSeedHash;
  if (k == 0) { nlive[slot] = 1; nframes[slot] = 0; }
}

__global__ void pool_gather_kernel(int B, int K, const int32_t* __restrict__ slots, const int32_t* __restrict__ ctx,
                                   const float* __restrict__ lp, const int32_t* __restrict__ len,
                                   const unsigned long long* __restrict__ hash, const int32_t* __restrict__ nlive,
                                   const int32_t* __restrict__ cst, BeamStateView d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int b = i / K, k = i - b * K;
  const size_t o = (size_t)slots[b] * K + k;
  if (d.cst != nullptr) d.cst[i] = cst[o];
  d.ctx[2 * i] = ctx[2 * o]; d.ctx[2 * i + 1] = ctx[2 * o + 1];
  d.lp[i] = lp[o]; d.len[i] = len[o]; d.hash[i] = hash[o];
  if (k == 0) d.nlive[b] = nlive[slots[b]];
}

// state back into the slots + the chunk's back-pointer rows appended to each stream's history
__global__ void pool_scatter_kernel(int B, int K, int Tc, int max_frames, const int32_t* __restrict__ slots, BeamStateView s,
                                    const int32_t* __restrict__ bp, int32_t* __restrict__ ctx, float* __restrict__ lp,
                                    int32_t* __restrict__ len, unsigned long long* __restrict__ hash, int32_t* __restrict__ nlive,
                                    int32_t* __restrict__ nframes, int32_t* __restrict__ hist, int32_t* __restrict__ cst) {
  const int b = blockIdx.x;
  const int slot = slots[b];
  const int f0 = nframes[slot];
  for (int i = threadIdx.x; i < Tc * K; i += blockDim.x)
    hist[((size_t)slot * max_frames + f0) * K + i] = bp[(size_t)b * Tc * K + i];
  if (threadIdx.x < K) {
    const size_t i = (size_t)b * K + threadIdx.x, o = (size_t)slot * K + threadIdx.x;
    ctx[2 * o] = s.ctx[2 * i]; ctx[2 * o + 1] = s.ctx[2 * i + 1];
    lp[o] = s.lp[i]; len[o] = s.len[i]; hash[o] = s.hash[i];
    if (s.cst != nullptr) cst[o] = s.cst[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) { nlive[slot] = s.nlive[b]; nframes[slot] = f0 + Tc; }
}

// One warp per stream: best hypothesis by log_prob / length (first maximum in slot order, as beam_backtrace_kernel), then the
// history is walked backwards through shared-memory tiles (the rows are fetched by the whole warp, the walk is lane 0's), twice:
// once to count the symbols, once to write them in forward order.
constexpr int kTileFrames = 256;
__global__ void __launch_bounds__(32)
pool_backtrace_kernel(int B, int K, int max_frames, const int32_t* __restrict__ slots, const float* __restrict__ lp,
                      const int32_t* __restrict__ len, const int32_t* __restrict__ nlive, const int32_t* __restrict__ nframes,
                      const int32_t* __restrict__ ctx, const int32_t* __restrict__ hist, int64_t* __restrict__ tokens,
                      int32_t* __restrict__ ts, int32_t* __restrict__ n_out, float* __restrict__ score, int64_t* __restrict__ hyp_out,
                      int cap, const int32_t* __restrict__ cst, const float* __restrict__ cg_resid) {
  __shared__ int32_t tile[kTileFrames * kMaxBeam];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int slot = slots[b];
  const int nf = nframes[slot], nl = nlive[slot];
  const int32_t* h = hist + (size_t)slot * max_frames * K;
  int best = 0;
  if (lane == 0) {
    float bn = -INFINITY;
    for (int q = 0; q < nl; ++q) {
      // (the boost of a hot word still being matched stays in the score while the stream is alive: only a finished utterance
      // revokes it - streaming results are "so far")
      const float norm = __fdiv_rn(lp[(size_t)slot * K + q], (float)len[(size_t)slot * K + q]);
      if (q == 0 || norm > bn) { bn = norm; best = q; }
    }
    (void)cst; (void)cg_resid;
    score[b] = lp[(size_t)slot * K + best];
    if (hyp_out != nullptr) {             // OnlineStream.Hyp <- last ctx tokens of the best hypothesis (ref OnlineRecognizer.cs:208)
      hyp_out[2 * b] = ctx[2 * ((size_t)slot * K + best)];
      hyp_out[2 * b + 1] = ctx[2 * ((size_t)slot * K + best) + 1];
    }
  }
  int n = 0;
  for (int pass = 0; pass < 2; ++pass) {
    int cur = best, i = 0;
    for (int hi = nf; hi > 0; hi -= kTileFrames) {
      const int lo = hi > kTileFrames ? hi - kTileFrames : 0;
      __syncwarp();
      for (int e = lane; e < (hi - lo) * K; e += 32) tile[e] = h[(size_t)lo * K + e];
      __syncwarp();
      if (lane == 0) {
        for (int t = hi - 1; t >= lo; --t) {
          const int e = tile[(t - lo) * K + cur];
          const int tok = (e & 0x0fffffff) - 1;
          if (tok >= 0) {
            if (pass == 1) {
              const int pos = n - 1 - i;
              if (pos < cap) { tokens[(size_t)b * cap + pos] = tok; ts[(size_t)b * cap + pos] = t; }
            }
            ++i;
          }
          cur = (e >> 28) & 0xf;
        }
      }
    }
    if (pass == 0) { n = i; if (lane == 0) n_out[b] = n; }
  }
}
}  // namespace

void beam_pool_free(k2b_handle* h) {
  BeamPool* p = h->beam_pool;
  if (p == nullptr) return;
  void* bufs[] = {p->ctx, p->lp, p->len, p->hash, p->nlive, p->nframes, p->hist, p->slots, p->cst};
  for (void* b : bufs) if (b) cudaFree(b);
  delete p;
  h->beam_pool = nullptr;
}

int beam_pool_K(const k2b_handle* h) { return h->beam_pool ? h->beam_pool->K : 0; }

int32_t beam_pool_create(k2b_handle* h, int max_streams, int K, int max_frames) {
  K2B_CUDA(h, cudaStreamSynchronize(h->stream));
  beam_pool_free(h);
  BeamPool* p = new BeamPool();
  h->beam_pool = p;
  p->K = K; p->max_streams = max_streams; p->max_frames = max_frames;
  const size_t S = (size_t)max_streams, N = S * K;
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->ctx), N * 2 * sizeof(int32_t)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->lp), N * sizeof(float)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->len), N * sizeof(int32_t)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->hash), N * sizeof(unsigned long long)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->cst), N * sizeof(int32_t)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->nlive), S * sizeof(int32_t)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->nframes), S * sizeof(int32_t)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->hist), N * (size_t)max_frames * sizeof(int32_t)));
  K2B_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&p->slots), S * sizeof(int32_t)));
  p->nframes_host.assign(S, 0);
  for (int s = 0; s < max_streams; ++s) K2B_TRY(beam_pool_reset(h, s, nullptr));
  return K2B_OK;
}

int32_t beam_pool_reset(k2b_handle* h, int slot, const int64_t* hyp_host) {
  BeamPool* p = h->beam_pool;
  if (p == nullptr) return fail(h, K2B_ERR_STATE, "k2b_beam_pool_reset: no pool (call k2b_beam_pool_create first)");
  if (slot < 0 || slot >= p->max_streams) return fail(h, K2B_ERR_INVALID, "k2b_beam_pool_reset: slot out of range");
  int c0 = -1, c1 = h->cfg.blank_id;
  if (hyp_host != nullptr) {
    const int V = h->cfg.vocab_size;
    if (hyp_host[0] < -1 || hyp_host[0] >= V || hyp_host[1] < 0 || hyp_host[1] >= V)
      return fail(h, K2B_ERR_INVALID, "k2b_beam_pool_reset: Hyp holds a token id outside the vocabulary");
    c0 = (int)hyp_host[0]; c1 = (int)hyp_host[1];
  }
  pool_reset_kernel<<<1, 32, 0, h->stream>>>(p->K, slot, c0, c1, h->cfg.blank_id, p->ctx, p->lp, p->len, p->hash, p->nlive, p->nframes, p->cst);
  K2B_LAUNCH_CHECK(h);
  p->nframes_host[(size_t)slot] = 0;
  return K2B_OK;
}

int32_t beam_pool_begin(k2b_handle* h, const int32_t* slots_host, int B, int Tc) {
  BeamPool* p = h->beam_pool;
  if (p == nullptr) return fail(h, K2B_ERR_STATE, "streaming beam search: no pool (call k2b_beam_pool_create first)");
  if (B > p->max_streams) return fail(h, K2B_ERR_INVALID, "streaming beam search: more streams than pool slots");
  std::vector<char> seen((size_t)p->max_streams, 0);
  for (int b = 0; b < B; ++b) {
    const int s = slots_host[b];
    if (s < 0 || s >= p->max_streams) return fail(h, K2B_ERR_INVALID, "streaming beam search: slot out of range");
    if (seen[(size_t)s]) return fail(h, K2B_ERR_INVALID, "streaming beam search: a slot appears twice in one call");
    seen[(size_t)s] = 1;
    if (p->nframes_host[(size_t)s] + Tc > p->max_frames)
      return fail(h, K2B_ERR_INVALID, "streaming beam search: stream longer than the pool's max_frames (reset the slot at an endpoint)");
  }
  // pageable source: the copy is staged before the call returns, and ordered on the stream behind earlier users of `slots`
  K2B_CUDA(h, cudaMemcpyAsync(p->slots, slots_host, sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
  p->cur_slots.assign(slots_host, slots_host + B);
  return K2B_OK;
}

int32_t beam_pool_gather(k2b_handle* h, int B, const BeamStateView& dst) {
  BeamPool* p = h->beam_pool;
  pool_gather_kernel<<<(B * p->K + 127) / 128, 128, 0, h->stream>>>(B, p->K, p->slots, p->ctx, p->lp, p->len, p->hash, p->nlive, p->cst, dst);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

int32_t beam_pool_scatter(k2b_handle* h, int B, int Tc, const BeamStateView& src, const int32_t* bp_chunk) {
  BeamPool* p = h->beam_pool;
  pool_scatter_kernel<<<B, 128, 0, h->stream>>>(B, p->K, Tc, p->max_frames, p->slots, src, bp_chunk, p->ctx, p->lp, p->len, p->hash,
                                                p->nlive, p->nframes, p->hist, p->cst);
  K2B_LAUNCH_CHECK(h);
  for (int s : p->cur_slots) p->nframes_host[(size_t)s] += Tc;
  return K2B_OK;
}

int32_t beam_pool_backtrace(k2b_handle* h, int B, int64_t* tokens, int32_t* ts, int32_t* n_out, float* score, int64_t* hyp_out, int cap) {
  BeamPool* p = h->beam_pool;
  pool_backtrace_kernel<<<B, 32, 0, h->stream>>>(B, p->K, p->max_frames, p->slots, p->lp, p->len, p->nlive, p->nframes, p->ctx, p->hist,
                                                 tokens, ts, n_out, score, hyp_out, cap, p->cst, h->cg_resid);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

}  // namespace k2b
