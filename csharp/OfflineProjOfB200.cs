// IOfflineProj over libk2b200.so: takes the place of OfflineProjOfTransducer (ref OfflineProjOfTransducer.cs).
// Lives in the K2TransducerAsr assembly because the interface's methods are `internal` (ref IOfflineProj.cs:43-47).
// NOT COMPILED HERE (no .NET toolchain in the build image); the Python mirror k2transducerasr_b200/proj.py is what CI runs.
using System;
using System.Collections.Generic;
using K2TransducerAsr.Model;
using Microsoft.ML.OnnxRuntime;
using Microsoft.ML.OnnxRuntime.Tensors;

namespace K2TransducerAsr.B200
{
    internal class OfflineProjOfB200 : IOfflineProj, IDisposable
    {
        private IntPtr _h;
        private readonly Func<List<OfflineInputEntity>, int, EncoderOutputEntity> _encoder;   // the (unchanged) encoder network
        private OfflineCustomMetadata _customMetadata;

        public OfflineProjOfB200(OfflineCustomMetadata meta, B200Weights w, Func<List<OfflineInputEntity>, int, EncoderOutputEntity> encoder,
                                 int device = 0, int precision = NativeMethods.PREC_BF16X3)
        {
            _customMetadata = meta;
            _encoder = encoder;
            var cfg = new K2bConfig {
                struct_size = 64, device = device, vocab_size = meta.Vocab_size, joiner_dim = meta.Joiner_dim,
                decoder_dim = w.DecoderDim, encoder_dim = w.EncoderDim, context_size = meta.Context_size,
                blank_id = 0, sos_eos_id = 1, unk_id = 2, max_beam = 4, precision = precision };
            int st = NativeMethods.k2b_create(ref cfg, out _h);
            if (st != 0) throw new Exception("k2b_create failed: " + System.Runtime.InteropServices.Marshal.PtrToStringUTF8(NativeMethods.k2b_last_error(IntPtr.Zero)));
            NativeMethods.Check(_h, NativeMethods.k2b_load_weights(_h, w.Emb, w.ConvW, w.DecProjW, w.DecProjB, w.EncProjW, w.EncProjB, w.OutW, w.OutB), "k2b_load_weights");
        }

        // The interface exposes ORT sessions (ref IOfflineProj.cs:8-22); there are none on this path.
        public InferenceSession EncoderSession { get => null!; set { } }
        public InferenceSession DecoderSession { get => null!; set { } }
        public InferenceSession JoinerSession { get => null!; set { } }
        public OfflineCustomMetadata CustomMetadata { get => _customMetadata; set => _customMetadata = value; }
        public int Blank_id { get; set; } = 0;
        public int Sos_eos_id { get; set; } = 1;
        public int Unk_id { get; set; } = 2;
        internal IntPtr Native => _h;

        public EncoderOutputEntity EncoderProj(List<OfflineInputEntity> modelInputs, int batchSize) => _encoder(modelInputs, batchSize);

        // ref OfflineProjOfTransducer.cs:93-123
        public DecoderOutputEntity DecoderProj(Int64[]? decoder_input, int batchSize)
        {
            int n = decoder_input == null ? batchSize : decoder_input.Length / _customMetadata.Context_size;
            var outp = new float[n * _customMetadata.Joiner_dim];
            NativeMethods.Check(_h, NativeMethods.k2b_decoder_proj(_h, decoder_input, n, outp), "DecoderProj");
            return new DecoderOutputEntity { decoder_out = outp };
        }

        // ref OfflineProjOfTransducer.cs:125-152
        public JoinerOutputEntity JoinerProj(float[]? encoder_out, float[]? decoder_out)
        {
            int J = _customMetadata.Joiner_dim, V = _customMetadata.Vocab_size, n = encoder_out!.Length / J;
            var logits = new float[n * V];
            NativeMethods.Check(_h, NativeMethods.k2b_joiner_proj(_h, encoder_out, decoder_out!, n, logits), "JoinerProj");
            return new JoinerOutputEntity { Logit = logits, Logits = new DenseTensor<float>(logits, new[] { n, V }, false) };
        }

        public void Dispose()
        {
            if (_h != IntPtr.Zero) { NativeMethods.k2b_destroy(_h); _h = IntPtr.Zero; }
            GC.SuppressFinalize(this);
        }
        ~OfflineProjOfB200() { if (_h != IntPtr.Zero) NativeMethods.k2b_destroy(_h); }
    }
}
