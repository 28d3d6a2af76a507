// k4lz4.hpp -- header-only C++ mirror of the reference's public block API over the C ABI
// (include/k4lz4.h).  Same names, argument meaning and error behaviour as
//   /root/reference/src/K4os.Compression.LZ4/LZ4Codec.cs:10-266, LZ4Level.cs:6-39,
//   LZ4Pickler.pickle.cs:51-106, LZ4Pickler.unpickle.cs:39-129
// for the accelerated path (L00_FAST encode, decode, byte[]-variant pickler).  The reference is
// compiled managed code; with no .NET toolchain in the build image this is the compiled-language
// host side above the C ABI (INTEGRATION.md shows the C# P/Invoke stubs).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "k4lz4.h"

namespace k4lz4 {

enum class LZ4Level : int {   // LZ4Level.cs:6-39
    L00_FAST = 0, L03_HC = 3, L04_HC = 4, L05_HC = 5, L06_HC = 6, L07_HC = 7, L08_HC = 8,
    L09_HC = 9, L10_OPT = 10, L11_OPT = 11, L12_MAX = 12
};

struct InvalidDataException : std::runtime_error { using std::runtime_error::runtime_error; };
struct NativeError : std::runtime_error {
    int code;
    NativeError(int c, const char* m) : std::runtime_error(std::string("libk4lz4: ") + m), code(c) {}
};
struct DelegateToManagedEngine : std::logic_error { using std::logic_error::logic_error; };

inline int check_codec(int r) {
    if (r <= K4LZ4_E_NODEVICE) throw NativeError(r, k4lz4_last_error());
    return r;
}

struct LZ4Codec {
    static constexpr int Version = 192;                                            // LZ4Codec.cs:13
    static int MaximumOutputSize(int length) { return k4lz4_max_output_size(length); }   // :30-31

    // LZ4Codec.Encode(byte*,int,byte*,int,LZ4Level) -- LZ4Codec.cs:40-52
    static int Encode(const uint8_t* source, int sourceLength, uint8_t* target, int targetLength,
                      LZ4Level level = LZ4Level::L00_FAST) {
        if (sourceLength <= 0) return 0;
        const int r = check_codec(k4lz4_encode(source, sourceLength, target, targetLength, (int)level));
        if (r == K4LZ4_R_DELEGATE) throw DelegateToManagedEngine("HC/OPT levels stay with the managed engine");
        return r;
    }
    // LZ4Codec.Decode(byte*,int,byte*,int) -- LZ4Codec.cs:104-115
    static int Decode(const uint8_t* source, int sourceLength, uint8_t* target, int targetLength) {
        if (sourceLength <= 0) return 0;
        return check_codec(k4lz4_decode(source, sourceLength, target, targetLength));
    }
};

struct LZ4Pickler {
    // LZ4Pickler.Pickle(ReadOnlySpan<byte>, LZ4Level) -- LZ4Pickler.pickle.cs:51-74
    static std::vector<uint8_t> Pickle(const uint8_t* source, int length, LZ4Level level = LZ4Level::L00_FAST) {
        if (length == 0) return {};
        std::vector<uint8_t> out((size_t)k4lz4_pickle_bound(length));
        int64_t so = 0, dof = 0; int32_t n = length, r = -1;
        const int rc = k4lz4_pickle_batch(source, &so, &n, out.data(), &dof, &r, 1, (int)level, K4LZ4_MEM_HOST, nullptr, 0);
        if (rc != K4LZ4_OK) throw NativeError(rc, k4lz4_last_error());
        if (r == K4LZ4_R_DELEGATE) throw DelegateToManagedEngine("HC/OPT levels stay with the managed engine");
        out.resize((size_t)r);
        return out;
    }
    // LZ4Pickler.UnpickledSize -- LZ4Pickler.unpickle.cs:83-92
    static int UnpickledSize(const uint8_t* source, int length) {
        int64_t so = 0; int32_t n = length, r = -1;
        const int rc = k4lz4_unpickled_size_batch(source, &so, &n, &r, 1, K4LZ4_MEM_HOST, nullptr, 0);
        if (rc != K4LZ4_OK) throw NativeError(rc, k4lz4_last_error());
        if (r == K4LZ4_R_CORRUPT) throw InvalidDataException("Pickle is corrupted");
        return r;
    }
    // LZ4Pickler.Unpickle(ReadOnlySpan<byte>) -- LZ4Pickler.unpickle.cs:39-50
    static std::vector<uint8_t> Unpickle(const uint8_t* source, int length) {
        if (length == 0) return {};
        const int size = UnpickledSize(source, length);
        std::vector<uint8_t> out((size_t)size);
        if (size == 0) return out;
        int64_t so = 0, dof = 0; int32_t n = length, dl = size, r = -1;
        const int rc = k4lz4_unpickle_batch(source, &so, &n, out.data(), &dof, &dl, &r, 1, K4LZ4_MEM_HOST, nullptr, 0);
        if (rc != K4LZ4_OK) throw NativeError(rc, k4lz4_last_error());
        if (r == K4LZ4_R_CORRUPT) throw InvalidDataException("Pickle is corrupted");
        return out;
    }
};

}  // namespace k4lz4
