// IOnlineProj over libk2b200.so: takes the place of OnlineProjOfZipformer / Zipformer2 / Lstm / Conformer on the search path
// (their DecoderProj / JoinerProj bodies are identical: ref OnlineProjOfZipformer2.cs:619-677 and siblings), and moves the
// per-chunk encoder-cache re-layout (ref OnlineProjOfZipformer2.cs:63-111, :144-489) onto the device.
// Lives in the K2TransducerAsr assembly (the interface's methods are `internal`, ref IOnlineProj.cs:65-71).
// NOT COMPILED HERE (no .NET toolchain); mirrored by k2transducerasr_b200/proj.py::OnlineProjOfB200 and tested on the GPU by
// tests/test_gpu_round2.py::test_online_recognizer_fused_on_gpu_with_device_states.
using System;
using System.Collections.Generic;
using System.Linq;
using K2TransducerAsr.Model;
using Microsoft.ML.OnnxRuntime;
using Microsoft.ML.OnnxRuntime.Tensors;

namespace K2TransducerAsr.B200
{
    /// The encoder network stays what it is (an ORT session, ideally on the CUDA EP with IOBinding): it receives the device address
    /// of the stacked caches and the chunk's features, and returns projected frames [B,T',J] (it has updated the caches in place).
    internal delegate EncoderOutputEntity OnlineEncoder(List<OnlineInputEntity> inputs, int batchSize, IntPtr stackedStatesDev, long stackedFloats);

    internal class OnlineProjOfB200 : IOnlineProj, IDisposable
    {
        private IntPtr _h;
        private readonly OnlineEncoder _encoder;
        private OnlineCustomMetadata _customMetadata;
        private readonly int[]? _itemLen, _axisLen;              // per-stream cache tensors and the "axisnum" of each (ref :236-341)
        private readonly Stack<int> _freeSlots = new();
        private IntPtr _stackedDev = IntPtr.Zero;                // device buffer of the call in flight (owned by the encoder side)
        private int[] _slotsInFlight = Array.Empty<int>();
        private readonly Func<long, IntPtr> _deviceAlloc;        // e.g. an OrtMemoryAllocation on the CUDA allocator

        public OnlineProjOfB200(OnlineCustomMetadata meta, B200Weights w, OnlineEncoder encoder, int[]? stateItemLen = null, int[]? stateAxisLen = null,
                                int maxStreams = 512, Func<long, IntPtr>? deviceAlloc = null, int device = 0, int precision = NativeMethods.PREC_BF16X3)
        {
            _customMetadata = meta; _encoder = encoder; _itemLen = stateItemLen; _axisLen = stateAxisLen; _deviceAlloc = deviceAlloc ?? (_ => IntPtr.Zero);
            var cfg = new K2bConfig {
                struct_size = 64, device = device, vocab_size = meta.Vocab_size, joiner_dim = meta.Joiner_dim, decoder_dim = w.DecoderDim,
                encoder_dim = w.EncoderDim, context_size = meta.Context_size, blank_id = 0, sos_eos_id = 1, unk_id = 2, max_beam = 8, precision = precision };
            if (NativeMethods.k2b_create(ref cfg, out _h) != 0)
                throw new Exception("k2b_create failed: " + System.Runtime.InteropServices.Marshal.PtrToStringUTF8(NativeMethods.k2b_last_error(IntPtr.Zero)));
            if (w.Emb != null)
                NativeMethods.Check(_h, NativeMethods.k2b_load_weights(_h, w.Emb, w.ConvW, w.DecProjW, w.DecProjB, w.EncProjW, w.EncProjB, w.OutW, w.OutB), "k2b_load_weights");
            if (_itemLen != null)
            {
                NativeMethods.Check(_h, NativeMethods.k2b_state_pool_create(_h, _itemLen, _itemLen.Length, maxStreams), "k2b_state_pool_create");
                for (int s = maxStreams - 1; s >= 0; s--) _freeSlots.Push(s);
            }
        }

        public InferenceSession EncoderSession { get => null!; set { } }      // ref IOnlineProj.cs:10-24: no ORT session on this path
        public InferenceSession DecoderSession { get => null!; set { } }
        public InferenceSession JoinerSession { get => null!; set { } }
        public OnlineCustomMetadata CustomMetadata { get => _customMetadata; set => _customMetadata = value; }
        public int Blank_id { get; set; } = 0;
        public int Sos_eos_id { get; set; } = 1;
        public int Unk_id { get; set; } = 2;
        public int ChunkLength { get; set; }
        public int ShiftLength { get; set; }
        public int FeatureDim { get; set; } = 80;
        public int SampleRate { get; set; } = 16000;
        internal IntPtr Native => _h;

        // The interface types carry host arrays (List<List<float[]>>); on this path a stream's caches never leave the GPU, so what
        // travels through OnlineStream.States is a one-element marker holding the stream's slot in the device pool.
        private static List<List<float[]>> Marker(int slot) => new() { new List<float[]> { new float[] { slot } } };
        private static int SlotOf(List<List<float[]>> states) => (int)states[0][0][0];

        // ref OnlineProjOfZipformer2.cs:63-111: all-zero caches of one stream -> a fresh slot, zero-filled on the device
        public List<List<float[]>> GetEncoderInitStates(int batchSize = 1)
        {
            if (_itemLen == null) return new List<List<float[]>>();
            int slot = _freeSlots.Pop();
            NativeMethods.Check(_h, NativeMethods.k2b_state_pool_put(_h, slot, new float[_itemLen.Sum()]), "k2b_state_pool_put");
            return Marker(slot);
        }
        public void ReleaseStates(List<List<float[]>> states) { if (_itemLen != null && states.Count > 0) _freeSlots.Push(SlotOf(states)); }

        // ref OnlineProjOfZipformer2.cs:144-362: per-stream caches -> batched tensors (batch on axis 1). One launch, no PCIe.
        public List<List<float[]>> stack_states(List<List<List<float[]>>> stateList)
        {
            if (_itemLen == null) return new List<List<float[]>>();
            _slotsInFlight = stateList.Select(SlotOf).ToArray();
            long n = NativeMethods.k2b_state_pool_stacked_floats(_h, _slotsInFlight.Length);
            _stackedDev = _deviceAlloc(n * sizeof(float));
            NativeMethods.Check(_h, NativeMethods.k2b_stack_states(_h, _slotsInFlight, _slotsInFlight.Length, _axisLen!, _stackedDev), "k2b_stack_states");
            NativeMethods.Check(_h, NativeMethods.k2b_sync(_h), "k2b_sync");
            return new List<List<float[]>>();                     // the batched caches are at _stackedDev
        }

        // ref OnlineProjOfZipformer2.cs:363-489: the encoder's new batched caches -> back into every stream's slot
        public List<List<List<float[]>>> unstack_states(List<float[]> encoder_out_states)
        {
            if (_itemLen == null) return new List<List<List<float[]>>>();
            NativeMethods.Check(_h, NativeMethods.k2b_unstack_states(_h, _slotsInFlight, _slotsInFlight.Length, _axisLen!, _stackedDev), "k2b_unstack_states");
            NativeMethods.Check(_h, NativeMethods.k2b_sync(_h), "k2b_sync");
            return _slotsInFlight.Select(Marker).ToList();
        }

        // ref OnlineProjOfZipformer2.cs:491-618
        public EncoderOutputEntity EncoderProj(List<OnlineInputEntity> modelInputs, int batchSize, List<List<float[]>> statesList)
        {
            try { return _encoder(modelInputs, batchSize, _stackedDev, NativeMethods.k2b_state_pool_stacked_floats(_h, batchSize)); }
            catch (Exception ex) { throw new Exception("EncoderProj failed", ex); }
        }

        // ref OnlineProjOfZipformer2.cs:619-648
        public virtual DecoderOutputEntity DecoderProj(Int64[]? decoder_input, int batchSize)
        {
            int n = decoder_input == null ? batchSize : decoder_input.Length / _customMetadata.Context_size;
            var outp = new float[n * _customMetadata.Joiner_dim];
            NativeMethods.Check(_h, NativeMethods.k2b_decoder_proj(_h, decoder_input, n, outp), "DecoderProj");
            return new DecoderOutputEntity { decoder_out = outp };
        }

        // ref OnlineProjOfZipformer2.cs:650-677
        public virtual JoinerOutputEntity JoinerProj(float[]? encoder_out, float[]? decoder_out)
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
        ~OnlineProjOfB200() { if (_h != IntPtr.Zero) NativeMethods.k2b_destroy(_h); }
    }
}
