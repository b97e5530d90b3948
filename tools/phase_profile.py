"""Per-phase cycle breakdown of the persistent cluster kernel on cfg2 (run on the GPU box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, synth, build

build.build()
cfg = synth.CONFIGS["cfg2"]
d = cfg.dims
names = ["enc prefetch issue (after the quarters)", "wait MMA", "TMEM->smem transpose + barrier", "reduce (keys, rounds, sum)",
         "st.async exchange + arrive", "wait partials (xbar)", "merge: warp 0 select_stream", "end-of-step barrier",
         "  sel: record loads", "  sel: lse + scoring", "  sel: K rounds", "  sel: extension + prefetch", "  sel: dedupe", "  sel: log-add",
         "  sel: write-back", "  build: quarter 0 (incl. load wait)", "  build: quarter 1", "  build: quarter 2", "  build: quarter 3"]
if __name__ == "__main__":
    # usage: phase_profile.py [prec:wh_tmem_kb[:blank_bias] ...]   (default: both precisions with the library's own placement;
    # a blank bias of e.g. 30 makes every frame blank: no new contexts, every decoder row an L2 hit)
    runs = [tuple(a.split(":")) for a in sys.argv[1:]] or [("bf16x3", "-1"), ("bf16", "-1")]
    for run in runs:
        prec, nt = run[0], run[1]
        bias = float(run[2]) if len(run) > 2 else cfg.blank_bias
        h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                           precision=_native.PREC_NAMES[prec])
        h.load_weights(synth.make_weights(d, blank_bias=bias))
        h.set_option("wh_tmem_kb", int(nt))
        h.set_option("pipe_chunks", 1)      # one kernel launch for the whole utterance (the host call would cut it into time chunks)
        raw = synth.make_frames(cfg.streams, cfg.frames, d.encoder_dim, cfg.seed)
        enc = h.encoder_proj(raw)           # projected frames in -> the host call is not time-chunked: one kernel launch
        h.modified_beam_search(enc, 4, enc_is_raw=False)
        h.cluster_phase_cycles()            # switch collection on
        h.modified_beam_search(enc, 4, enc_is_raw=False)
        cyc = h.cluster_phase_cycles()
        tot = cyc[:8].sum() + cyc[15:19].sum()
        print(f"== {prec} wh_tmem_kb={nt} blank_bias={bias}: {tot / cfg.frames:.0f} cycles per frame step (CTA 0)")
        for n, c in zip(names, cyc[:19]):
            print(f"   {n:32s} {c / cfg.frames:8.0f} cyc  {100.0 * c / tot:5.1f} %")
        h.close()
