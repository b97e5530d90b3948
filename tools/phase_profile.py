"""Per-phase cycle breakdown of the persistent cluster kernel on cfg2 (run on the GPU box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build

build.build()
cfg = synth.CONFIGS["cfg2"]
d = cfg.dims
names = ["build x (gather+tanh+split)", "sync+fence", "MMA issue", "MMA wait", "TMEM->smem transpose", "reduce+DSMEM st",
         "cluster barrier", "(after merge) sync wait", "  reduce: loads+max+sum+sort+rounds", "  merge: warp 0 select_stream", "-", "-",
         "    select: lse", "    select: candidate scoring", "    select: K rounds", "    select: extension+dedupe", "    select: log-add", "    select: write-back", "-", "-"]
for prec in ("bf16x3", "bf16"):
    h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                       precision=_native.PREC_NAMES[prec])
    h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
    raw = synth.make_frames(cfg.streams, cfg.frames, d.encoder_dim, cfg.seed)
    enc = h.encoder_proj(raw)           # projected frames in -> the host call is not time-chunked: one kernel launch
    h.modified_beam_search(enc, 4, enc_is_raw=False)
    h.cluster_phase_cycles()            # switch collection on
    h.modified_beam_search(enc, 4, enc_is_raw=False)
    cyc = h.cluster_phase_cycles()
    tot = cyc.sum()
    print(f"== {prec}: {tot / cfg.frames:.0f} cycles per frame step (CTA 0)")
    for n, c in zip(names, cyc):
        print(f"   {n:32s} {c / cfg.frames:8.0f} cyc  {100.0 * c / tot:5.1f} %")
    h.close()
