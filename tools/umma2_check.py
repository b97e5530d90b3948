"""CTA-pair (cta_group::2) tcgen05.mma self-test against numpy (run on the GPU box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, build


def round_bf16(x):
    """float32 -> nearest-even bfloat16 -> float32"""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32)


build.build()
h = _native.Handle(vocab_size=64, joiner_dim=64, decoder_dim=64)
rng = np.random.default_rng(5)
for ts in (False, True):
    for N, K in ((64, 64), (32, 128), (128, 256), (256, 64)):
        A = rng.standard_normal((256, K), dtype=np.float32); B = rng.standard_normal((N, K), dtype=np.float32)
        D = h.selftest_umma2(A, B, ts)
        want = round_bf16(A).astype(np.float64) @ round_bf16(B).astype(np.float64).T
        err = np.abs(D - want).max()
        print(f"ts={int(ts)} N={N:3d} K={K:3d}: max |err| {err:.3e}  {'OK' if err < 1e-3 else 'MISMATCH'}", flush=True)
h.close()
