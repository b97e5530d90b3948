"""cfg2 modified_beam_search from PAGEABLE host memory (a plain numpy array, as a managed float[] is), wall clock per call for a
sweep of the library's copy-thread counts; the page-locked call beside it. Prints the host's thread count too."""
import os
import sys
import time
sys.path.insert(0, ".")
import numpy as np
import torch
from k2transducerasr_b200 import _native, synth, build

build.build()
cfg = synth.CONFIGS["cfg2"]
d = cfg.dims
h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                   precision=_native.PREC_NAMES["bf16x3"])
h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
raws = [synth.make_frames(cfg.streams, cfg.frames, d.encoder_dim, cfg.seed + i) for i in range(2)]
pinned = [torch.from_numpy(r).pin_memory().numpy() for r in raws]
print("host threads:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))


def run(src, n=6):
    h.modified_beam_search(src[0], cfg.beam)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(n):
        h.modified_beam_search(src[i % 2], cfg.beam)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


print("page-locked (includes the Python unpacking of the results): %.3f ms" % run(pinned))
for th in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "-1,1,2,4,6,8,12,16".split(","))]:
    h.set_option("copy_threads", th)
    print("copy_threads %3d: %.3f ms" % (th, run(raws)))
h.close()
