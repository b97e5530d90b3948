import sys
sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, build
h = _native.Handle(vocab_size=64, joiner_dim=64, decoder_dim=64)
print(h.selftest_collectives())
