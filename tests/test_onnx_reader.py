"""CPU: the dependency-free ONNX initialiser / custom-metadata reader (k2transducerasr_b200/onnx_reader.py) on the hand-built
fixtures under tests/golden/ - what stands where the reference opens its three InferenceSessions and reads their metadata map
(ref OfflineModel.cs:31-72, :84-118)."""
from pathlib import Path

import numpy as np
import pytest

from k2transducerasr_b200 import onnx_reader as R

G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def truth():
    return np.load(G / "tiny_onnx_weights.npz")


def test_weights_and_metadata_from_onnx(truth):
    w = R.B200Weights.FromOnnx(G / "tiny_decoder.onnx", G / "tiny_joiner.onnx", G / "tiny_encoder.onnx")
    assert (w.context_size, w.vocab_size, w.joiner_dim, w.decoder_dim, w.encoder_dim) == (2, 11, 16, 16, 32)
    assert w.metadata["model_type"] == "zipformer2" and w.metadata["comment"] == "k2b200 fixture"
    for k in ("emb", "conv_w", "dec_proj_w", "dec_proj_b", "out_w", "out_b", "enc_proj_w", "enc_proj_b"):
        np.testing.assert_array_equal(w.weights[k], truth[k], err_msg=k)      # Gemm(transB) as is, MatMul transposed, biases found
        assert w.weights[k].dtype == np.float32 and w.weights[k].flags["C_CONTIGUOUS"]
    no_enc = R.B200Weights.FromOnnx(G / "tiny_decoder.onnx", G / "tiny_joiner.onnx")
    assert no_enc.encoder_dim == 0 and no_enc.weights["enc_proj_w"] is None


def test_int8_joiner_is_dequantised(truth):
    j, meta = R.read_joiner(G / "tiny_joiner.int8.onnx")
    want = (truth["q"].astype(np.float32) * truth["scale"]).T
    np.testing.assert_array_equal(j["out_w"], want)
    np.testing.assert_array_equal(j["out_b"], truth["out_b"])
    assert np.abs(j["out_w"] - truth["out_w"]).max() <= truth["scale"] * 0.5 + 1e-7
    assert meta["joiner_dim"] == "16"


def test_wire_format_details(tmp_path):
    # unpacked repeated dims, negative int64 data, a scalar, fp16 raw data, an attribute - and an external-data tensor is refused
    t = R._ld(8, b"v") + R._vi(1, 2) + R._vi(1, 3) + R._vi(2, 7) + b"".join(R._vi(7, x) for x in (-1, 2, -3, 4, 5, 6))
    m = R.enc_model([R.enc_node("Gemm", ["a", "v"], ["o"], {"transB": 1, "alpha": 3})], [t, R.enc_tensor("h", np.arange(4, dtype=np.float16)),
                                                                                       R.enc_tensor("s", np.asarray(2.5, np.float32))], {"k": "v"})
    p = tmp_path / "m.onnx"
    p.write_bytes(m)
    mod = R.load(p)
    assert mod.initializers["v"].tolist() == [[-1, 2, -3], [4, 5, 6]] and mod.initializers["v"].dtype == np.int64
    assert mod.initializers["h"].dtype == np.float16 and mod.initializers["s"].shape == () and float(mod.initializers["s"]) == 2.5
    assert mod.nodes[0].op_type == "Gemm" and mod.nodes[0].ints == {"transB": 1, "alpha": 3} and mod.metadata == {"k": "v"}
    ext = R._ld(8, b"x") + R._vi(1, 4) + R._vi(2, 1) + R._vi(14, 1)
    p.write_bytes(R.enc_model([], [ext], {}))
    with pytest.raises(ValueError, match="external"):
        R.load(p)
    p.write_bytes(R.enc_model([R.enc_node("Relu", ["a"], ["b"])], [], {}))
    with pytest.raises(ValueError):
        R.read_decoder(p)
