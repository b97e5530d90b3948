"""Online greedy chunk time against the chunk length on cfg3's shape (512 streams, V = 2000): slope = per-frame step, intercept =
per-call fixed cost. K2B_GREEDY_PERSISTENT=0/1 picks the engine. Run on the GPU box."""
import sys
import torch
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, build, synth
build.build()
cfg = synth.CONFIGS["cfg3"]; d = cfg.dims
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES["bf16x3"])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias)); h.set_stream(stream.cuda_stream)
B = cfg.streams
for Tc in (4, 8, 16, 32, 64):
    x = torch.from_numpy(synth.make_frames(B, Tc, d.encoder_dim, 5)).to(dev)
    hyp = torch.zeros((B, 2), dtype=torch.int64, device=dev)
    tok = torch.zeros((B, Tc), dtype=torch.int64, device=dev); ts = torch.zeros((B, Tc), dtype=torch.int32, device=dev)
    n = torch.zeros((B,), dtype=torch.int32, device=dev)
    f = lambda: h.call("k2b_greedy_online_chunk_dev", x, 1, B, Tc, hyp, tok, ts, n, Tc)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20): f()
    e1.record(stream); torch.cuda.synchronize()
    print(f"Tc={Tc:3d}: {1e3 * e0.elapsed_time(e1) / 20:8.1f} us per chunk")
