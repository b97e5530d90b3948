"""cfg4 per-frame beam path: CUDA-event time of each of the three launches of a frame step (operand build, joiner, merge) and the
joiner's in-kernel phase cycles (CTA 0: start -> accumulator ready, TMEM -> shared tile, reduction). Run on the GPU box."""
import os
import subprocess
import sys

sys.path.insert(0, ".")

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from k2transducerasr_b200 import _native, synth, build
    build.build()
    cfg = synth.CONFIGS["cfg4"]; d = cfg.dims
    h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                       precision=_native.PREC_NAMES["bf16x3"])
    h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
    T = 16
    raw = synth.make_frames(cfg.streams, T, d.encoder_dim, cfg.seed)
    h.modified_beam_search(raw, 4, enc_is_raw=True)
    which = int(os.environ.get("K2B_PROF_WHICH", "0"))
    if which == 0:
        h.cluster_phase_cycles()
    h.profile_enable(True)
    h.modified_beam_search(raw, 4, enc_is_raw=True)
    n, ms = h.profile_read()
    print(f"which={which} ({['joiner', 'operand build', 'merge'][which]}): {n} launches, avg {1e3 * ms / max(n, 1):.2f} us")
    if which == 0:
        c = h.cluster_phase_cycles()
        k = max(int(c[13]), 1)
        print(f"  joiner CTA 0 cycles per launch: start->acc {c[10] / k:.0f}, tmem->tile {c[11] / k:.0f}, reduce {c[12] / k:.0f} ({k} launches)")
    import time
    import torch
    t0 = time.perf_counter()
    for _ in range(3):
        h.modified_beam_search(raw, 4, enc_is_raw=True)
    dt = (time.perf_counter() - t0) / 3
    print(f"  whole call (host pointers, T={T}): {1e6 * dt / T:.1f} us per frame")
else:
    for w in ("0", "1", "2"):
        env = dict(os.environ, K2B_PROF_WHICH=w)
        subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
