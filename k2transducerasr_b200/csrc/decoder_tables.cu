// Stateless-decoder front half, done once at weight-load time.
//
// The decoder inside decoder.onnx (ref OfflineProjOfTransducer.cs:116; math [EXT] icefall) is
//   e[k] = Emb[y[k]] (zero row for a masked negative id), k < ctx
//   c[o] = sum_k sum_{i<4} conv_w[o,i,k] * e[k][4*(o/4)+i]          Conv1d(groups = D/4, no bias)
//   out  = dec_proj(relu(c))
// The grouped conv is linear in each context position, so it is folded into one [V+1,D] table per
// position:  tab_k[v][o] = sum_i conv_w[o,i,k] * Emb[v][4*(o/4)+i]   and   c = tab_0[y0] + tab_1[y1].
// Per search step the decoder front half is then a pure two-row gather + add + ReLU (HBM/L2-bound),
// fused into the prologue of the decoder_proj GEMM. Row V of each table is zero (masked id).
#include "k2b_internal.h"

namespace k2b {

namespace {

__global__ void fold_conv_kernel(const float* __restrict__ emb, const float* __restrict__ conv_w,
                                 float* __restrict__ tab0, float* __restrict__ tab1, int V, int D, int ctx) {
  const int v = blockIdx.x;
  for (int o = threadIdx.x; o < D; o += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    if (v < V) {
      const float* e = emb + (size_t)v * D + 4 * (o / 4);
      const float* w = conv_w + (size_t)o * 4 * ctx;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s0 = fmaf(w[i * ctx + 0], e[i], s0);
        s1 = fmaf(w[i * ctx + 1], e[i], s1);
      }
    }
    tab0[(size_t)v * D + o] = s0;
    tab1[(size_t)v * D + o] = s1;
  }
}

}  // namespace

int32_t build_decoder_tables(k2b_handle* h) {
  const int V = h->cfg.vocab_size, D = h->cfg.decoder_dim, ctx = h->cfg.context_size;
  fold_conv_kernel<<<V + 1, 128, 0, h->stream>>>(h->emb, h->conv_w, h->tab0, h->tab1, V, D, ctx);
  K2B_LAUNCH_CHECK(h);
  return K2B_OK;
}

}  // namespace k2b
