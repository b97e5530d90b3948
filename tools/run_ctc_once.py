"""One large CTC greedy call + one tcgen05 encoder_proj call (for ncu captures)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build
build.build()
dev = torch.device("cuda", 0)
cfg = synth.CONFIGS["cfg2"]; d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_BF16X3)
h.load_weights(synth.make_weights(d))
B, T, V = 1024, 250, 2000
logp = torch.log_softmax(torch.randn((B, T, V), device=dev) * 3, -1).contiguous()
tok = torch.zeros((B, T), dtype=torch.int64, device=dev); ts = torch.zeros((B, T), dtype=torch.int32, device=dev)
n = torch.zeros((B,), dtype=torch.int32, device=dev)
raw = torch.randn((64000, d.encoder_dim), device=dev)
out = torch.empty((64000, d.joiner_dim), device=dev)
for _ in range(2):
    h.call("k2b_ctc_greedy_dev", logp, B, T, V, 0, None, None, tok, ts, n, None, T)
    h.call("k2b_encoder_proj_dev", raw, 64000, out)
h.sync()
print("ok", int(n.sum().item()))
