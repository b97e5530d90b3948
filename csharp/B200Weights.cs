// B200Weights.FromOnnx: the weights k2b_load_weights takes, read straight from the reference's model files
// (decoder*.onnx / joiner*.onnx, optionally the encoder_proj tail of encoder*.onnx) without ONNX Runtime: an ONNX file is a
// protobuf ModelProto whose initialisers are plain TensorProtos. Same walker as k2transducerasr_b200/onnx_reader.py, which CI
// runs against tests/golden/tiny_*.onnx. The custom-metadata map is the one ref OfflineModel.cs:31-72 reads.
// NOT COMPILED HERE (no .NET toolchain in the build image).
using System;
using System.Collections.Generic;
using System.IO;
using System.Linq;
using System.Text;

namespace K2TransducerAsr.B200
{
    internal class B200Weights
    {
        public int ContextSize, VocabSize, JoinerDim, DecoderDim, EncoderDim;
        public float[] Emb = null!, ConvW = null!, DecProjW = null!, DecProjB = null!, OutW = null!, OutB = null!;
        public float[]? EncProjW, EncProjB;
        public Dictionary<string, string> Metadata = new();

        public static B200Weights FromOnnx(string decoderFilePath, string joinerFilePath, string? encoderFilePath = null)
        {
            var dec = OnnxFile.Load(decoderFilePath);
            var joi = OnnxFile.Load(joinerFilePath);
            var w = new B200Weights();
            foreach (var kv in joi.Metadata) w.Metadata[kv.Key] = kv.Value;
            foreach (var kv in dec.Metadata) w.Metadata[kv.Key] = kv.Value;
            var emb = dec.FirstInput("Gather", 0) ?? throw new InvalidDataException("decoder.onnx: no Gather with a constant table");
            var conv = dec.FirstInput("Conv", 1) ?? throw new InvalidDataException("decoder.onnx: no Conv with a constant weight");
            var dl = dec.Linears().LastOrDefault() ?? throw new InvalidDataException("decoder.onnx: no decoder_proj Linear");
            var jl = joi.Linears().LastOrDefault() ?? throw new InvalidDataException("joiner.onnx: no output Linear");
            w.VocabSize = (int)emb.Dims[0]; w.DecoderDim = (int)emb.Dims[1]; w.JoinerDim = jl.In;
            w.ContextSize = w.Metadata.TryGetValue("context_size", out var cs) ? int.Parse(cs) : (int)conv.Dims[2];   // ref OfflineModel.cs:34-35
            w.Emb = emb.Data; w.ConvW = conv.Data; w.DecProjW = dl.W; w.DecProjB = dl.B ?? new float[dl.Out];
            w.OutW = jl.W; w.OutB = jl.B ?? new float[jl.Out];
            if (jl.Out != w.VocabSize || dl.Out != w.JoinerDim || dl.In != w.DecoderDim)
                throw new InvalidDataException("decoder.onnx / joiner.onnx disagree on vocab_size / joiner_dim / decoder_dim");
            if (encoderFilePath != null)
            {
                var el = OnnxFile.Load(encoderFilePath).Linears().LastOrDefault(l => l.Out == w.JoinerDim)
                         ?? throw new InvalidDataException("encoder.onnx: no Linear with joiner_dim outputs");
                w.EncProjW = el.W; w.EncProjB = el.B ?? new float[el.Out]; w.EncoderDim = el.In;
            }
            return w;
        }
    }

    internal sealed class OnnxTensor { public string Name = ""; public long[] Dims = Array.Empty<long>(); public int Type = 1; public float[] Data = Array.Empty<float>(); }
    internal sealed class OnnxNode { public string Op = ""; public List<string> In = new(), Out = new(); public Dictionary<string, long> Ints = new(); }
    internal sealed class Linear { public float[] W = null!; public float[]? B; public int Out, In; }

    /// Wire-format walker: ModelProto{7 graph, 14 metadata_props}, GraphProto{1 node, 5 initializer},
    /// NodeProto{1 input, 2 output, 4 op_type, 5 attribute{1 name, 3 i}}, TensorProto{1 dims, 2 data_type, 4 float_data, 8 name, 9 raw_data}.
    internal sealed class OnnxFile
    {
        public Dictionary<string, string> Metadata = new();
        public Dictionary<string, OnnxTensor> Init = new();
        public List<OnnxNode> Nodes = new();

        private static ulong Varint(byte[] b, ref int p) { ulong v = 0; int s = 0; while (true) { byte x = b[p++]; v |= (ulong)(x & 0x7F) << s; if ((x & 0x80) == 0) return v; s += 7; } }

        private static IEnumerable<(int num, int wt, ulong v, int off, int len)> Fields(byte[] b, int start, int end)
        {
            int p = start;
            while (p < end)
            {
                ulong key = Varint(b, ref p); int num = (int)(key >> 3), wt = (int)(key & 7);
                if (wt == 0) { ulong v = Varint(b, ref p); yield return (num, wt, v, 0, 0); }
                else if (wt == 1) { yield return (num, wt, 0, p, 8); p += 8; }
                else if (wt == 5) { yield return (num, wt, 0, p, 4); p += 4; }
                else if (wt == 2) { int n = (int)Varint(b, ref p); yield return (num, wt, 0, p, n); p += n; }
                else throw new InvalidDataException("unsupported protobuf wire type " + wt);
            }
        }

        public static OnnxFile Load(string path)
        {
            var b = File.ReadAllBytes(path);
            var f = new OnnxFile();
            foreach (var (num, _, _, off, len) in Fields(b, 0, b.Length))
            {
                if (num == 14)
                {
                    string k = "", v = "";
                    foreach (var (n2, _, _, o2, l2) in Fields(b, off, off + len)) { if (n2 == 1) k = Encoding.UTF8.GetString(b, o2, l2); else if (n2 == 2) v = Encoding.UTF8.GetString(b, o2, l2); }
                    f.Metadata[k] = v;
                }
                else if (num == 7)
                    foreach (var (n2, _, _, o2, l2) in Fields(b, off, off + len))
                    {
                        if (n2 == 5) { var t = Tensor(b, o2, o2 + l2); f.Init[t.Name] = t; }
                        else if (n2 == 1) f.Nodes.Add(Node(b, o2, o2 + l2));
                    }
            }
            return f;
        }

        private static OnnxTensor Tensor(byte[] b, int start, int end)
        {
            var t = new OnnxTensor(); var dims = new List<long>(); var fl = new List<float>(); (int off, int len) raw = (0, -1);
            foreach (var (num, wt, v, off, len) in Fields(b, start, end))
            {
                if (num == 1) { if (wt == 0) dims.Add((long)v); else { int p = off; while (p < off + len) dims.Add((long)Varint(b, ref p)); } }
                else if (num == 2) t.Type = (int)v;
                else if (num == 4) { for (int i = 0; i < len; i += 4) fl.Add(BitConverter.ToSingle(b, off + i)); }
                else if (num == 8) t.Name = Encoding.UTF8.GetString(b, off, len);
                else if (num == 9) raw = (off, len);
                else if (num == 14 && v == 1) throw new InvalidDataException($"tensor {t.Name} keeps its data in an external file");
            }
            t.Dims = dims.ToArray();
            if (raw.len >= 0)
            {
                if (t.Type == 1) { t.Data = new float[raw.len / 4]; Buffer.BlockCopy(b, raw.off, t.Data, 0, raw.len); }
                else if (t.Type == 3) t.Data = Enumerable.Range(0, raw.len).Select(i => (float)(sbyte)b[raw.off + i]).ToArray();     // int8 (quantised exports)
                else if (t.Type == 2) t.Data = Enumerable.Range(0, raw.len).Select(i => (float)b[raw.off + i]).ToArray();            // uint8
                else if (t.Type == 10) t.Data = Enumerable.Range(0, raw.len / 2).Select(i => (float)BitConverter.ToHalf(b, raw.off + 2 * i)).ToArray();
                else throw new InvalidDataException($"tensor {t.Name}: ONNX data type {t.Type} is not supported");
            }
            else t.Data = fl.ToArray();
            return t;
        }

        private static OnnxNode Node(byte[] b, int start, int end)
        {
            var n = new OnnxNode();
            foreach (var (num, _, _, off, len) in Fields(b, start, end))
            {
                if (num == 1) n.In.Add(Encoding.UTF8.GetString(b, off, len));
                else if (num == 2) n.Out.Add(Encoding.UTF8.GetString(b, off, len));
                else if (num == 4) n.Op = Encoding.UTF8.GetString(b, off, len);
                else if (num == 5)
                {
                    string an = ""; long? iv = null;
                    foreach (var (n2, _, v2, o2, l2) in Fields(b, off, off + len)) { if (n2 == 1) an = Encoding.UTF8.GetString(b, o2, l2); else if (n2 == 3) iv = (long)v2; }
                    if (iv.HasValue) n.Ints[an] = iv.Value;
                }
            }
            return n;
        }

        /// fp32 view of an initialiser; `<x>_quantized` is de-quantised with `<x>_scale` / `<x>_zero_point` (onnxruntime quantiser naming).
        private OnnxTensor? Fp32(string name)
        {
            if (!Init.TryGetValue(name, out var t)) return null;
            if (t.Type != 2 && t.Type != 3) return t;
            string bas = name.EndsWith("_quantized") ? name.Substring(0, name.Length - 10) : name;
            if (!Init.TryGetValue(bas + "_scale", out var sc)) throw new InvalidDataException($"quantised tensor {name} has no {bas}_scale");
            float zp = Init.TryGetValue(bas + "_zero_point", out var z) && z.Data.Length > 0 ? z.Data[0] : 0f;
            return new OnnxTensor { Name = name, Dims = t.Dims, Type = 1, Data = t.Data.Select(x => (x - zp) * sc.Data[0]).ToArray() };
        }

        public OnnxTensor? FirstInput(string op, int index) => Nodes.Where(n => n.Op == op && n.In.Count > index).Select(n => Fp32(n.In[index])).FirstOrDefault(t => t != null);

        private static float[] Transpose(float[] a, int rows, int cols) { var o = new float[a.Length]; for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) o[c * rows + r] = a[r * cols + c]; return o; }

        /// Every Linear in node order as [out,in] weight + bias: Gemm honours transB; MatMul / MatMulInteger weights are [in,out] and the
        /// bias is the constant of the next Add on the path.
        public IEnumerable<Linear> Linears()
        {
            for (int i = 0; i < Nodes.Count; i++)
            {
                var nd = Nodes[i];
                if (nd.Op == "Gemm")
                {
                    var w = Fp32(nd.In[1]); if (w == null) continue;
                    bool tb = nd.Ints.TryGetValue("transB", out var t) && t != 0;
                    int r = (int)w.Dims[0], c = (int)w.Dims[1];
                    yield return new Linear { W = tb ? w.Data : Transpose(w.Data, r, c), Out = tb ? r : c, In = tb ? c : r, B = nd.In.Count > 2 ? Fp32(nd.In[2])?.Data : null };
                }
                else if (nd.Op == "MatMul" || nd.Op == "MatMulInteger")
                {
                    var w = Fp32(nd.In[1]); if (w == null || w.Dims.Length != 2) continue;
                    int r = (int)w.Dims[0], c = (int)w.Dims[1];
                    float[]? bias = null; var frontier = new HashSet<string>(nd.Out);
                    foreach (var nx in Nodes.Skip(i + 1).Take(7))
                    {
                        if (!nx.In.Any(frontier.Contains)) continue;
                        if (nx.Op == "Add") { bias = nx.In.Select(Fp32).FirstOrDefault(x => x != null && x.Data.Length == c)?.Data; break; }
                        foreach (var o in nx.Out) frontier.Add(o);
                    }
                    yield return new Linear { W = Transpose(w.Data, r, c), Out = c, In = r, B = bias };
                }
            }
        }
    }
}
