"""encoder_proj on the tensor cores alone (k2b_encoder_proj_dev): cfg2's 64000 x 768 -> 512 in one launch and in 32-frame chunks'
worth of rows, cfg3's 4096 rows; CUDA events on the launch stream."""
import sys
sys.path.insert(0, ".")
import torch
from k2transducerasr_b200 import _native, synth, build

build.build()
cfg = synth.CONFIGS["cfg2"]
d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES["bf16x3"])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
h.set_stream(stream.cuda_stream)
for rows in (64000, 8192, 6400, 4096):
    x = torch.randn(rows, d.encoder_dim, device="cuda")
    y = torch.empty(rows, d.joiner_dim, device="cuda")
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    h.call("k2b_encoder_proj_dev", x, rows, y)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h.call("k2b_encoder_proj_dev", x, rows, y)
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    flop = 2.0 * rows * d.encoder_dim * d.joiner_dim * 3
    print("rows %6d: median %7.1f us  min %7.1f us  (%.0f TFLOP/s of split-bf16 x3 at the median)" % (rows, ts[len(ts) // 2], ts[0], flop / ts[len(ts) // 2] * 1e-6))
# the last batch against fp64 torch (split-bf16 x3 is fp32-grade)
w = synth.make_weights(d, blank_bias=cfg.blank_bias)
ref = (x.double() @ torch.from_numpy(w["enc_proj_w"]).cuda().double().T + torch.from_numpy(w["enc_proj_b"]).cuda().double())
print("max abs error against fp64: %.3g" % float((y.double() - ref).abs().max()))
h.close()
