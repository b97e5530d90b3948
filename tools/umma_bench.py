import sys
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, build
build.build()
h = _native.Handle(vocab_size=64, joiner_dim=64, decoder_dim=64)
names = ["SS M128 N32", "SS M128 N64", "TS M128 N32", "TS M128 N64", "SS M64  N32", "SS M128 N128", "TS M128 N128",
         "SS N32 2 accumulators", "SS N32 4 accumulators", "SS N32 8 accumulators", "SS N64 4 accumulators",
         "warp-uniform SS N32", "warp-uniform TS N32", "warp-uniform SS N64",
         "wu SS N64 2 accumulators", "wu SS N64 4 accumulators", "wu pair SS64+TS32 same acc (x2 MMAs)", "wu pair separate acc (x2)", "wu pair separate acc, K over 2 (x2)"]
for f, nm in enumerate(names):
    if f < 11: continue
    for reps in (1, 8):
        h.selftest_umma_bench(f, 4, reps)
        c = h.selftest_umma_bench(f, 4, reps)
        n = reps * 16
        print(f"{nm}: {n:4d} MMAs  issue {c[0] / n:7.1f} cyc/MMA   complete {c[1] / n:7.1f} cyc/MMA  (total {c[1]})")
h.close()
