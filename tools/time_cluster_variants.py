"""cfg2 cluster kernel with 0 .. 8 k-blocks of W_hi in tensor memory (k2b_set_option "wh_tmem_kb"): device time per launch (CUDA
events on the launch stream, one launch per utterance) and equality of the outputs with the all-in-shared-memory placement."""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build

build.build()
cfg = synth.CONFIGS["cfg2"]
d = cfg.dims
precs = sys.argv[1:] or ["bf16x3", "bf16"]
for prec in precs:
    h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                       precision=_native.PREC_NAMES[prec])
    h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
    raw = synth.make_frames(cfg.streams, cfg.frames, d.encoder_dim, cfg.seed)
    enc = h.encoder_proj(raw)
    base = None
    h.set_option("pipe_chunks", 1)
    for nt in (0, -1):      # 0 = everything in shared memory, -1 = the library's choice (6 of 8 k-blocks / all 8 in tensor memory)
        h.set_option("wh_tmem_kb", nt)
        try:
            for _ in range(2):
                out = h.modified_beam_search(enc, cfg.beam, enc_is_raw=False)
        except _native.K2bError as e:
            print(f"{prec} nt={nt}: {e}")
            continue
        h.profile_enable(True)
        n = 10
        for _ in range(n):
            out = h.modified_beam_search(enc, cfg.beam, enc_is_raw=False)
        nl, ms = h.profile_read()
        h.profile_enable(False)
        us = ms * 1e3 / nl
        if base is None:
            base = out
        same = sum(1 for a, b in zip(out[0], base[0]) if a == b)
        dsc = float(np.max(np.abs(np.asarray(out[2]) - np.asarray(base[2]))))
        print(f"{prec} wh_tmem_kb={nt}: {us:8.1f} us per launch, {us / cfg.frames * 1.965e3:7.0f} cycles per frame step; "
              f"streams identical to wh_tmem_kb=0: {same}/{cfg.streams}, max score diff {dsc:.2e}", flush=True)
    h.close()
