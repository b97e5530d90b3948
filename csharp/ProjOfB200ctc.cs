// CTC variants of the seam over libk2b200.so (ref OfflineProjOfZipformer2ctc.cs, OnlineProjOfZipformer2ctc.cs): EncoderProj returns
// [B,T,V] log-probs from the (unchanged) encoder network, DecoderProj / JoinerProj return null (ref OfflineProjOfZipformer2ctc.cs:93-101,
// OnlineProjOfZipformer2ctc.cs:616-624); the search itself is FusedSearch.Forward*GreedySearchCTC* -> k2b_ctc_greedy, which needs no
// weights. NOT COMPILED HERE (no .NET toolchain); mirrored by proj.py::OfflineProjOfB200ctc / OnlineProjOfB200ctc and tested by
// tests/test_gpu_round2.py::test_ctc_projs_through_the_recognizers.
using System;
using System.Collections.Generic;
using K2TransducerAsr.Model;
using Microsoft.ML.OnnxRuntime;

namespace K2TransducerAsr.B200
{
    internal class OfflineProjOfB200ctc : IOfflineProj, IDisposable
    {
        private IntPtr _h;
        private readonly Func<List<OfflineInputEntity>, int, EncoderOutputEntity> _encoder;
        private OfflineCustomMetadata _customMetadata;

        public OfflineProjOfB200ctc(OfflineCustomMetadata meta, Func<List<OfflineInputEntity>, int, EncoderOutputEntity> encoder, int device = 0)
        {
            _customMetadata = meta; _encoder = encoder;
            // k2b_ctc_greedy takes V as an argument (the reference takes it from tokens.txt, ref OfflineRecognizer.cs:325) and needs no weights
            var cfg = new K2bConfig { struct_size = 64, device = device, vocab_size = Math.Max(meta.Vocab_size, 1), joiner_dim = 64, decoder_dim = 64,
                                      encoder_dim = 0, context_size = 2, blank_id = 0, sos_eos_id = 1, unk_id = 2, max_beam = 1 };
            if (NativeMethods.k2b_create(ref cfg, out _h) != 0)
                throw new Exception("k2b_create failed: " + System.Runtime.InteropServices.Marshal.PtrToStringUTF8(NativeMethods.k2b_last_error(IntPtr.Zero)));
        }
        public InferenceSession EncoderSession { get => null!; set { } }
        public InferenceSession DecoderSession { get => null!; set { } }
        public InferenceSession JoinerSession { get => null!; set { } }
        public OfflineCustomMetadata CustomMetadata { get => _customMetadata; set => _customMetadata = value; }
        public int Blank_id { get; set; } = 0;
        public int Sos_eos_id { get; set; } = 1;
        public int Unk_id { get; set; } = 2;
        internal IntPtr Native => _h;
        public EncoderOutputEntity EncoderProj(List<OfflineInputEntity> modelInputs, int batchSize) => _encoder(modelInputs, batchSize);
        public DecoderOutputEntity DecoderProj(Int64[]? decoder_input, int batchSize) => null!;      // ref OfflineProjOfZipformer2ctc.cs:93-96
        public JoinerOutputEntity JoinerProj(float[]? encoder_out, float[]? decoder_out) => null!;  // ref :98-101
        public void Dispose() { if (_h != IntPtr.Zero) { NativeMethods.k2b_destroy(_h); _h = IntPtr.Zero; } GC.SuppressFinalize(this); }
        ~OfflineProjOfB200ctc() { if (_h != IntPtr.Zero) NativeMethods.k2b_destroy(_h); }
    }

    /// Online: the cache handling of OnlineProjOfB200 (device pool) with the two null bodies (ref OnlineProjOfZipformer2ctc.cs:616-624 -
    /// that file is otherwise a copy of OnlineProjOfZipformer2.cs, SURVEY.md section 2 #4). `new B200Weights()` (no tensors) is enough.
    internal class OnlineProjOfB200ctc : OnlineProjOfB200
    {
        public OnlineProjOfB200ctc(OnlineCustomMetadata meta, OnlineEncoder encoder, int[]? stateItemLen = null, int[]? stateAxisLen = null,
                                   int maxStreams = 512, Func<long, IntPtr>? deviceAlloc = null, int device = 0)
            : base(meta, new B200Weights { DecoderDim = 64, EncoderDim = 0 }, encoder, stateItemLen, stateAxisLen, maxStreams, deviceAlloc, device, NativeMethods.PREC_FP32) { }
        public override DecoderOutputEntity DecoderProj(Int64[]? decoder_input, int batchSize) => null!;
        public override JoinerOutputEntity JoinerProj(float[]? encoder_out, float[]? decoder_out) => null!;
    }
}
