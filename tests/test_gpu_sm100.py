"""GPU (-m gpu): hardware self-tests of the tcgen05 / TMEM / bulk-TMA / cluster building blocks, against numpy.
These pin the descriptor and layout encodings of csrc/sm100_ptx.cuh on a real B200."""
import numpy as np
import pytest

from k2transducerasr_b200 import _native
from oracle.k2_oracle import round_bf16

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h(built_lib):
    hd = _native.Handle(vocab_size=64, joiner_dim=64, decoder_dim=64)
    yield hd
    hd.close()


@pytest.mark.parametrize("K,N", [(64, 16), (256, 32), (128, 32)])
@pytest.mark.parametrize("mode", [0, 3])
def test_umma_bf16_operands(h, K, N, mode):
    rng = np.random.default_rng(K + N + mode)
    A = rng.standard_normal((128, K), dtype=np.float32)
    B = rng.standard_normal((N, K), dtype=np.float32)
    D = h.selftest_umma(A, B, mode)
    want = round_bf16(A).astype(np.float64) @ round_bf16(B).astype(np.float64).T
    np.testing.assert_allclose(D, want, rtol=0, atol=2e-4)


@pytest.mark.parametrize("mode", [1, 2])
def test_umma_split_bf16x3_is_fp32_accurate(h, mode):
    rng = np.random.default_rng(7 + mode)
    A = rng.standard_normal((128, 256), dtype=np.float32)
    B = rng.standard_normal((32, 256), dtype=np.float32)
    D = h.selftest_umma(A, B, mode)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    assert np.abs(D - want).max() < 4e-4          # |a.b| ~ 16; bf16x3 drops only the lo*lo term (2^-16 relative)
    assert np.abs(h.selftest_umma(A, B, 0) - want).max() > 10 * np.abs(D - want).max()


def test_umma_operand_through_bulk_tma(h):
    rng = np.random.default_rng(99)
    A = rng.standard_normal((128, 256), dtype=np.float32)
    B = rng.standard_normal((32, 256), dtype=np.float32)
    np.testing.assert_array_equal(h.selftest_umma(A, B, 0, use_tma=True), h.selftest_umma(A, B, 0, use_tma=False))


@pytest.mark.parametrize("csize,ncl", [(1, 3), (2, 5), (4, 37), (8, 18), (16, 4)])
def test_cluster_dsmem_roundtrip(h, csize, ncl):
    bad, done = h.selftest_cluster(csize, ncl)
    assert bad == 0 and done == csize * ncl
