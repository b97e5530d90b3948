"""Writes tests/golden/tiny_{decoder,joiner,encoder}.onnx (+ tiny_onnx_weights.npz with the tensors that went in): hand-built ONNX
protobufs of the three graphs the reference loads (ref OfflineModel.cs:84-118), a few hundred bytes each, in the two shapes real
exports take - Gemm(transB=1) with raw_data initialisers, and MatMul + Add with packed float_data - plus a dynamically quantised
joiner (MatMulInteger, the onnxruntime quantiser's `_quantized` / `_scale` / `_zero_point` naming) and the custom metadata map.
    python tests/golden/make_onnx_fixtures.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from k2transducerasr_b200 import onnx_reader as R     # noqa: E402
from k2transducerasr_b200 import synth                 # noqa: E402

OUT = Path(__file__).resolve().parent
DIMS = synth.ModelDims(vocab_size=11, joiner_dim=16, decoder_dim=16, encoder_dim=32)


def main():
    w = synth.make_weights(DIMS, seed=3)
    meta = {"context_size": "2", "vocab_size": "11", "model_type": "zipformer2", "version": "1", "comment": "k2b200 fixture"}
    # decoder.onnx: Gather -> (mask etc. elided) -> Conv -> Relu -> Gemm(transB = 1); initialisers renamed as exporters do
    nodes = [R.enc_node("Gather", ["onnx::Gather_7", "y"], ["e"]), R.enc_node("Conv", ["e", "decoder.conv.weight"], ["c"], {"group": 4}),
             R.enc_node("Relu", ["c"], ["r"]), R.enc_node("Gemm", ["r", "decoder_proj.weight", "decoder_proj.bias"], ["decoder_out"], {"transB": 1})]
    tens = [R.enc_tensor("onnx::Gather_7", w["emb"]), R.enc_tensor("decoder.conv.weight", w["conv_w"]),
            R.enc_tensor("decoder_proj.weight", w["dec_proj_w"]), R.enc_tensor("decoder_proj.bias", w["dec_proj_b"])]
    (OUT / "tiny_decoder.onnx").write_bytes(R.enc_model(nodes, tens, meta))
    # joiner.onnx: Add -> Tanh -> MatMul ([in,out] weight, packed float_data) -> Add(bias)
    nodes = [R.enc_node("Add", ["encoder_out", "decoder_out"], ["s"]), R.enc_node("Tanh", ["s"], ["t"]),
             R.enc_node("MatMul", ["t", "onnx::MatMul_21"], ["mm"]), R.enc_node("Add", ["mm", "output_linear.bias"], ["logit"])]
    tens = [R.enc_tensor("onnx::MatMul_21", np.ascontiguousarray(w["out_w"].T), raw=False), R.enc_tensor("output_linear.bias", w["out_b"], raw=False)]
    (OUT / "tiny_joiner.onnx").write_bytes(R.enc_model(nodes, tens, {"joiner_dim": "16"}))
    # joiner.int8.onnx: DynamicQuantizeLinear -> MatMulInteger -> Cast -> Mul -> Add
    wt = np.ascontiguousarray(w["out_w"].T)
    scale = np.float32(np.abs(wt).max() / 127.0)
    q = np.clip(np.round(wt / scale), -127, 127).astype(np.int8)
    nodes = [R.enc_node("Add", ["encoder_out", "decoder_out"], ["s"]), R.enc_node("Tanh", ["s"], ["t"]),
             R.enc_node("DynamicQuantizeLinear", ["t"], ["tq", "ts", "tz"]),
             R.enc_node("MatMulInteger", ["tq", "W_quantized", "tz", "W_zero_point"], ["mi"]), R.enc_node("Cast", ["mi"], ["mf"], {"to": 1}),
             R.enc_node("Mul", ["ts", "W_scale"], ["sc"]), R.enc_node("Mul", ["mf", "sc"], ["mm"]),
             R.enc_node("Add", ["mm", "output_linear.bias"], ["logit"])]
    tens = [R.enc_tensor("W_quantized", q), R.enc_tensor("W_scale", np.asarray(scale, np.float32)), R.enc_tensor("W_zero_point", np.asarray(0, np.int8)),
            R.enc_tensor("output_linear.bias", w["out_b"])]
    (OUT / "tiny_joiner.int8.onnx").write_bytes(R.enc_model(nodes, tens, {"joiner_dim": "16"}))
    # encoder.onnx tail: ... -> MatMul -> Add (encoder_proj), preceded by an unrelated Linear of another width
    other = np.random.default_rng(1).standard_normal((32, 32)).astype(np.float32)
    nodes = [R.enc_node("MatMul", ["x", "ff.weight"], ["h"]), R.enc_node("MatMul", ["h", "onnx::MatMul_99"], ["p"]),
             R.enc_node("Add", ["p", "encoder_proj.bias"], ["encoder_out"])]
    tens = [R.enc_tensor("ff.weight", other), R.enc_tensor("onnx::MatMul_99", np.ascontiguousarray(w["enc_proj_w"].T)),
            R.enc_tensor("encoder_proj.bias", w["enc_proj_b"])]
    (OUT / "tiny_encoder.onnx").write_bytes(R.enc_model(nodes, tens, {"model_type": "zipformer2", "decode_chunk_len": "32", "T": "45"}))
    np.savez_compressed(OUT / "tiny_onnx_weights.npz", q=q, scale=scale, **{k: v for k, v in w.items() if v is not None})
    print("wrote", sorted(p.name for p in OUT.glob("tiny_*")))


if __name__ == "__main__":
    main()
