// k4lz4.hpp -- header-only C++ mirror of the reference's public block API over the C ABI
// (include/k4lz4.h).  Same names, argument meaning and error behaviour as
//   /root/reference/src/K4os.Compression.LZ4/LZ4Codec.cs:10-266, LZ4Level.cs:6-39,
//   LZ4Pickler.pickle.cs:51-106, LZ4Pickler.unpickle.cs:39-129
// for the accelerated path (L00_FAST encode, decode incl. dictionary / partial, byte[]-variant
// pickler) and Encoders/LZ4BlockEncoder.cs, LZ4BlockDecoder.cs with a batched top-up.  The reference is
// compiled managed code; with no .NET toolchain in the build image this is the compiled-language
// host side above the C ABI (INTEGRATION.md shows the C# P/Invoke stubs).
#pragma once
#include <algorithm>
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
    // LZ4Codec.Decode(byte*,int,byte*,int,byte*,int) -- LZ4Codec.cs:144-157
    static int Decode(const uint8_t* source, int sourceLength, uint8_t* target, int targetLength,
                      const uint8_t* dictionary, int dictionaryLength) {
        if (sourceLength <= 0) return 0;
        return check_codec(k4lz4_decode_dict(source, sourceLength, target, targetLength, dictionary, dictionaryLength));
    }
    // LZ4Codec.PartialDecode(byte*,int,byte*,int) -- LZ4Codec.cs:123-134
    static int PartialDecode(const uint8_t* source, int sourceLength, uint8_t* target, int targetLength) {
        if (sourceLength <= 0) return 0;
        return check_codec(k4lz4_partial_decode(source, sourceLength, target, targetLength));
    }
};

// Independent-block stream pair with a batched top-up:
//   Encoders/LZ4EncoderBase.cs:27-97, Encoders/LZ4BlockEncoder.cs:7-23, Encoders/LZ4BlockDecoder.cs:11-102.
// Same members as the reference's ILZ4Encoder / ILZ4Decoder for one block at a time, plus
// TopupMany / EncodeMany / DecodeMany, which move a whole queue of blocks through ONE
// k4lz4_encode_batch / k4lz4_decode_batch call.  Chained (dependent-block) encoders are not data
// parallel and stay with the managed engine.
class LZ4BlockEncoder {
public:
    struct Block { int encoded; std::vector<uint8_t> bytes; };   // encoded < 0: stored raw (allowCopy)

    LZ4BlockEncoder(LZ4Level level, int blockSize, int batchBlocks = 256)
        : level_(level), block_(roundUp(blockSize < 1024 ? 1024 : blockSize, 1024)),
          depth_(batchBlocks < 1 ? 1 : batchBlocks), buf_((size_t)depth_ * block_), fill_((size_t)depth_, 0) {}

    int BlockSize() const { return block_; }
    int BytesReady() const { return cur_ < depth_ ? fill_[(size_t)cur_] : 0; }
    int BlocksQueued() const { int n = 0; for (int f : fill_) n += f > 0; return n; }

    // LZ4EncoderBase.cs:47-62
    int Topup(const uint8_t* source, int length) {
        if (length <= 0 || cur_ >= depth_) return 0;
        const int left = block_ - fill_[(size_t)cur_];
        if (left <= 0) return 0;
        const int chunk = left < length ? left : length;
        std::copy(source, source + chunk, buf_.begin() + (size_t)cur_ * block_ + fill_[(size_t)cur_]);
        fill_[(size_t)cur_] += chunk;
        return chunk;
    }
    // fills block after block until the queue or the source is exhausted
    int64_t TopupMany(const uint8_t* source, int64_t length) {
        int64_t taken = 0;
        while (taken < length && cur_ < depth_) {
            const int64_t rest = length - taken;
            const int got = Topup(source + taken, (int)(rest > block_ ? block_ : rest));
            taken += got;
            if (fill_[(size_t)cur_] == block_) cur_++;
            else if (got == 0) break;
        }
        return taken;
    }
    // LZ4EncoderBase.cs:65-87 for every queued block, one GPU call
    std::vector<Block> EncodeMany(bool allowCopy = true) {
        const int nb = BlocksQueued();
        std::vector<Block> out;
        if (nb == 0) return out;
        if ((int)level_ >= 3) throw DelegateToManagedEngine("HC/OPT levels stay with the managed engine");
        const int bound = k4lz4_max_output_size(block_);
        std::vector<int64_t> so((size_t)nb), dof((size_t)nb);
        std::vector<int32_t> sl((size_t)nb), cap((size_t)nb, bound), res((size_t)nb, -1);
        for (int i = 0; i < nb; i++) { so[(size_t)i] = (int64_t)i * block_; dof[(size_t)i] = (int64_t)i * bound; sl[(size_t)i] = fill_[(size_t)i]; }
        std::vector<uint8_t> dst((size_t)nb * bound);
        const int rc = k4lz4_encode_batch(buf_.data(), so.data(), sl.data(), dst.data(), dof.data(), cap.data(),
                                          res.data(), nb, (int)level_, K4LZ4_MEM_HOST, nullptr, K4LZ4_ALL_DEVICES);
        if (rc != K4LZ4_OK) throw NativeError(rc, k4lz4_last_error());
        for (int i = 0; i < nb; i++) {
            const int enc = res[(size_t)i], n = sl[(size_t)i];
            if (enc <= 0) throw std::runtime_error("Failed to encode chunk. Target buffer too small.");
            Block b;
            if (allowCopy && enc >= n) { b.encoded = -n; b.bytes.assign(buf_.begin() + so[(size_t)i], buf_.begin() + so[(size_t)i] + n); }
            else { b.encoded = enc; b.bytes.assign(dst.begin() + dof[(size_t)i], dst.begin() + dof[(size_t)i] + enc); }
            out.push_back(std::move(b));
        }
        std::fill(fill_.begin(), fill_.end(), 0);        // Commit(): independent blocks keep no dictionary
        cur_ = 0;
        return out;
    }
    // the reference's single-block call
    int Encode(uint8_t* target, int length, bool allowCopy) {
        if (BlocksQueued() == 0) return 0;
        if (BlocksQueued() != 1) throw std::logic_error("Encode() handles one pending block; use EncodeMany()");
        const int n = fill_[0];
        int enc = LZ4Codec::Encode(buf_.data(), n, target, length, level_);
        if (enc <= 0) throw std::runtime_error("Failed to encode chunk. Target buffer too small.");
        if (allowCopy && enc >= n) { std::copy(buf_.begin(), buf_.begin() + n, target); enc = -n; }
        fill_[0] = 0; cur_ = 0;
        return enc;
    }

private:
    static int roundUp(int v, int step) { return (v + step - 1) / step * step; }
    LZ4Level level_;
    int block_, depth_, cur_ = 0;
    std::vector<uint8_t> buf_;
    std::vector<int> fill_;
};

class LZ4BlockDecoder {
public:
    explicit LZ4BlockDecoder(int blockSize)
        : block_(((blockSize < 1024 ? 1024 : blockSize) + 1023) / 1024 * 1024), outLen_(block_ + 8), out_((size_t)outLen_ + 8) {}
    int BlockSize() const { return block_; }
    int BytesReady() const { return index_; }
    // LZ4BlockDecoder.cs:39-55
    int Decode(const uint8_t* source, int length, int blockSize = 0) {
        if (blockSize <= 0) blockSize = block_;
        if (blockSize > block_) throw std::runtime_error("InvalidOperationException");
        const int decoded = LZ4Codec::Decode(source, length, out_.data(), outLen_);
        if (decoded < 0) throw std::runtime_error("InvalidOperationException");
        return index_ = decoded;
    }
    // one GPU call for a list of compressed blocks (raw = true: the encoder stored the block as is)
    struct Item { const uint8_t* data; int length; bool raw; };
    std::vector<std::vector<uint8_t>> DecodeMany(const std::vector<Item>& items) {
        const int n = (int)items.size();
        std::vector<std::vector<uint8_t>> res((size_t)n);
        if (n == 0) return res;
        std::vector<int64_t> so((size_t)n), dof((size_t)n);
        std::vector<int32_t> sl((size_t)n), cap((size_t)n, outLen_), got((size_t)n, -1);
        int64_t tot = 0;
        for (int i = 0; i < n; i++) { so[(size_t)i] = tot; sl[(size_t)i] = items[(size_t)i].raw ? 0 : items[(size_t)i].length; tot += sl[(size_t)i]; dof[(size_t)i] = (int64_t)i * outLen_; }
        std::vector<uint8_t> src((size_t)tot + 16), dst((size_t)n * outLen_ + 16);
        for (int i = 0; i < n; i++) if (sl[(size_t)i] > 0) std::copy(items[(size_t)i].data, items[(size_t)i].data + sl[(size_t)i], src.begin() + so[(size_t)i]);
        const int rc = k4lz4_decode_batch(src.data(), so.data(), sl.data(), dst.data(), dof.data(), cap.data(), got.data(),
                                          n, K4LZ4_MEM_HOST, nullptr, K4LZ4_ALL_DEVICES);
        if (rc != K4LZ4_OK) throw NativeError(rc, k4lz4_last_error());
        for (int i = 0; i < n; i++) {
            const Item& it = items[(size_t)i];
            if (it.raw) {
                if (it.length > outLen_) throw std::runtime_error("InvalidOperationException");
                res[(size_t)i].assign(it.data, it.data + it.length);
            } else {
                if (got[(size_t)i] < 0 || (got[(size_t)i] == 0 && it.length > 0)) throw std::runtime_error("InvalidOperationException");
                res[(size_t)i].assign(dst.begin() + dof[(size_t)i], dst.begin() + dof[(size_t)i] + got[(size_t)i]);
            }
        }
        std::copy(res.back().begin(), res.back().end(), out_.begin());
        index_ = (int)res.back().size();
        return res;
    }
    // LZ4BlockDecoder.cs:58-71
    int Inject(const uint8_t* source, int length) {
        if (length <= 0) return index_ = 0;
        if (length > outLen_) throw std::runtime_error("InvalidOperationException");
        std::copy(source, source + length, out_.begin());
        return index_ = length;
    }
    // LZ4BlockDecoder.cs:74-83 (offset is negative: counted from the end of the block)
    void Drain(uint8_t* target, int offset, int length) const {
        offset = index_ + offset;
        if (offset < 0 || length < 0 || offset + length > index_) throw std::runtime_error("InvalidOperationException");
        std::copy(out_.begin() + offset, out_.begin() + offset + length, target);
    }
    const uint8_t* Peek(int offset) const {
        offset = index_ + offset;
        if (offset < 0 || offset > index_) throw std::runtime_error("InvalidOperationException");
        return out_.data() + offset;
    }

private:
    int block_, outLen_, index_ = 0;
    std::vector<uint8_t> out_;
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
