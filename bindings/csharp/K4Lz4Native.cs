// K4Lz4Native.cs -- P/Invoke surface of libk4lz4 (include/k4lz4.h) for K4os.Compression.LZ4.
// Drop into src/K4os.Compression.LZ4/Engine/Native/ (see INTEGRATION.md for the two call sites
// in LZ4Codec.cs that change).  Not compiled in this repository: the build image has no .NET SDK.
using System;
using System.Runtime.InteropServices;

namespace K4os.Compression.LZ4.Engine.Native
{
    internal static unsafe class K4Lz4Native
    {
        private const string Lib = "k4lz4";

        public const int R_DELEGATE = -2;      // level not handled natively (HC/OPT stay managed)
        public const int R_CORRUPT = -1000;    // InvalidDataException ("Pickle is corrupted")
        public const int E_NODEVICE = -100;    // <= -100: library-level failure, see k4lz4_last_error()
        public const int MEM_HOST = 0, MEM_DEVICE = 1, ALL_DEVICES = -1;

        [DllImport(Lib)] public static extern int k4lz4_codec_version();
        [DllImport(Lib)] public static extern int k4lz4_device_count();
        [DllImport(Lib)] public static extern sbyte* k4lz4_last_error();
        [DllImport(Lib)] public static extern int k4lz4_max_output_size(int length);
        [DllImport(Lib)] public static extern int k4lz4_pickle_bound(int length);
        [DllImport(Lib)] public static extern int k4lz4_encode(byte* src, int srcLen, byte* dst, int dstCap, int level);
        [DllImport(Lib)] public static extern int k4lz4_decode(byte* src, int srcLen, byte* dst, int dstCap);
        [DllImport(Lib)] public static extern int k4lz4_encode_batch(
            byte* srcBase, long* srcOff, int* srcLen, byte* dstBase, long* dstOff, int* dstCap,
            int* outLen, int nBlocks, int level, int memKind, void* cudaStream, int device);
        [DllImport(Lib)] public static extern int k4lz4_decode_batch(
            byte* srcBase, long* srcOff, int* srcLen, byte* dstBase, long* dstOff, int* dstCap,
            int* outLen, int nBlocks, int memKind, void* cudaStream, int device);
        [DllImport(Lib)] public static extern int k4lz4_pickle_batch(
            byte* srcBase, long* srcOff, int* srcLen, byte* dstBase, long* dstOff,
            int* outLen, int nMessages, int level, int memKind, void* cudaStream, int device);
        [DllImport(Lib)] public static extern int k4lz4_unpickled_size_batch(
            byte* srcBase, long* srcOff, int* srcLen, int* outSize, int nMessages,
            int memKind, void* cudaStream, int device);
        [DllImport(Lib)] public static extern int k4lz4_unpickle_batch(
            byte* srcBase, long* srcOff, int* srcLen, byte* dstBase, long* dstOff, int* dstLen,
            int* outLen, int nMessages, int memKind, void* cudaStream, int device);

        // decode with an external dictionary / partial decode (LZ4Codec.cs:123-134,144-157)
        [DllImport(Lib)] public static extern int k4lz4_decode_dict(
            byte* src, int srcLen, byte* dst, int dstCap, byte* dict, int dictLen);
        [DllImport(Lib)] public static extern int k4lz4_partial_decode(byte* src, int srcLen, byte* dst, int targetLen);
        [DllImport(Lib)] public static extern int k4lz4_decode_dict_batch(
            byte* srcBase, long* srcOff, int* srcLen, byte* dstBase, long* dstOff, int* dstCap,
            byte* dictBase, long* dictOff, int* dictLen, int* outLen, int nBlocks,
            int memKind, void* cudaStream, int device);
        [DllImport(Lib)] public static extern int k4lz4_partial_decode_batch(
            byte* srcBase, long* srcOff, int* srcLen, byte* dstBase, long* dstOff, int* targetLen,
            int* outLen, int nBlocks, int memKind, void* cudaStream, int device);
        // LL.Enforce32 semantics (LL.tools.cs:29-36) for inputs >= 65 547 bytes
        [DllImport(Lib)] public static extern int k4lz4_encode_x32(byte* src, int srcLen, byte* dst, int dstCap, int level);
        [DllImport(Lib)] public static extern int k4lz4_encode_batch_x32(
            byte* srcBase, long* srcOff, int* srcLen, byte* dstBase, long* dstOff, int* dstCap,
            int* outLen, int nBlocks, int level, int memKind, void* cudaStream, int device);
        // Pickle<TBufferWriter> (LZ4Pickler.pickle.cs:113-148)
        [DllImport(Lib)] public static extern int k4lz4_pickle_writer_bound(int length);
        [DllImport(Lib)] public static extern int k4lz4_pickle_writer_batch(
            byte* srcBase, long* srcOff, int* srcLen, byte* dstBase, long* dstOff,
            int* outLen, int nMessages, int level, int memKind, void* cudaStream, int device);

        public static string LastError() => new string(k4lz4_last_error());
    }
}
