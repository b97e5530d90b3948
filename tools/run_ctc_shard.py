"""cfg5 at one GPU's share (128 streams x 250 frames x V=2000 = 256 MB per call, two alternating inputs): back-to-back launch timing
of the CTC kernel's two frame schedules, and the launches an ncu capture picks up. Run on the GPU box."""
import sys
import torch
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, build
build.build()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
h = _native.Handle(vocab_size=64, joiner_dim=64, decoder_dim=64)
h.set_stream(stream.cuda_stream)
B, T, V = (int(sys.argv[1]) if len(sys.argv) > 1 else 128), 250, 2000
xs = [torch.log_softmax(torch.randn((B, T, V), device=dev) * 3, -1).contiguous() for _ in range(2)]
tok = torch.zeros((B, T), dtype=torch.int64, device=dev); ts = torch.zeros((B, T), dtype=torch.int32, device=dev)
n = torch.zeros((B,), dtype=torch.int32, device=dev)
for static in (1, 0):        # k2b_set_option("ctc_one_kernel"): 1 = tickets + collapse inside one kernel, 0 = frames + collapse kernels (PDL)
    h.set_option("ctc_one_kernel", static)
    for i in range(4):
        h.call("k2b_ctc_greedy_dev", xs[i % 2], B, T, V, 0, None, None, tok, ts, n, None, T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(20):
        h.call("k2b_ctc_greedy_dev", xs[i % 2], B, T, V, 0, None, None, tok, ts, n, None, T)
    e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"B={B} ctc_one_kernel={static}: {us:.1f} us per launch back to back, {B * T * V * 4 / us / 1e3:.0f} GB/s, emitted {int(n.sum().item())}")
for name, fn in (("torch.sum", lambda x: x.sum()), ("torch.amax(-1)", lambda x: x.amax(-1))):
    for i in range(4):
        fn(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(20):
        fn(xs[i % 2])
    e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"B={B} {name}: {us:.1f} us per call back to back, {B * T * V * 4 / us / 1e3:.0f} GB/s  (a library read-reduction over the same bytes)")
h.close()
