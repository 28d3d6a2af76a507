"""LZ4 Frame container around BATCHED independent blocks.

Writer and reader follow the reference's frame code for the case this library accelerates --
independent blocks (`Chaining = false`), `L00_FAST`:
    Streams/Frames/LZ4FrameWriter.cs:57-108 (header: magic 0x184D2204, FLG, BD, HC),
    :159-189 (block length code with bit 31 = stored raw, XXH32 block / content checksums),
    LZ4FrameWriter.blocking.cs:22-33,88-97 (block = length code, data, [checksum]; tail = end mark,
    [content checksum]), LZ4FrameReader.blocking.cs:57-144 (header / block parsing and checks);
    format: orig/doc/lz4_Frame_format.md.
All blocks of a frame go through ONE k4lz4_encode_batch / k4lz4_decode_batch call and ONE
k4lz4_xxh32_batch call; only the serial parts (header byte, content checksum, byte layout) run
on the host.  Frames are interoperable with upstream lz4 (tests decode them with
orig/lib/lz4frame.c and decode upstream's frames here).  Chained blocks, content size and
dictionary ids are the managed engine's business (the reference itself throws NotImplemented for
the last two, LZ4FrameWriter.cs:89-95).
"""
from __future__ import annotations

import struct

import numpy as np

from . import _native as N
from .batch import decode_batch_flat_host, encode_batch_flat_host

MAGIC = 0x184D2204
_BLOCK_SIZES = {4: 1 << 16, 5: 1 << 18, 6: 1 << 20, 7: 1 << 22}


class InvalidDataException(ValueError):
    """Malformed frame (the reference throws InvalidDataException from LZ4FrameReader)."""


def xxh32(data, seed: int = 0) -> int:
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return int(N.lib().k4lz4_xxh32(a.ctypes.data if a.size else None, int(a.size), seed))


def xxh32_batch(base: np.ndarray, off, length, seed: int = 0, device: int = 0) -> np.ndarray:
    off = np.ascontiguousarray(off, dtype=np.int64)
    length = np.ascontiguousarray(length, dtype=np.int32)
    out = np.zeros(len(length), dtype=np.uint32)
    if len(length):
        N.check(N.lib().k4lz4_xxh32_batch(base.ctypes.data, off.ctypes.data, length.ctypes.data, seed,
                                          out.ctypes.data, len(length), N.MEM_HOST, None, device))
    return out


def _block_size_code(block_size: int) -> int:                # LZ4FrameWriter.cs:183-188
    for code in (4, 5, 6, 7):
        if block_size <= _BLOCK_SIZES[code]:
            return code
    raise ValueError(f"Invalid block size {block_size} for stream")


def write_frame(data, block_size: int = 65536, block_checksum: bool = False,
                content_checksum: bool = False, level: int = 0, device: int = 0) -> bytes:
    """One LZ4 frame of independent blocks holding `data` (LZ4FrameWriter with Chaining = false)."""
    src = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    code = _block_size_code(block_size)
    bs = max(1024, (block_size + 1023) // 1024 * 1024)          # LZ4EncoderBase.cs:29
    flg = (1 << 6) | (1 << 5) | (int(block_checksum) << 4) | (int(content_checksum) << 2)
    bd = code << 4
    head = struct.pack("<IBB", MAGIC, flg, bd)
    out = [head, bytes([(xxh32(head[4:6]) >> 8) & 0xFF])]       # HC, LZ4FrameWriter.cs:100-102
    n = int(src.size)
    nb = (n + bs - 1) // bs
    if nb:
        lens = np.full(nb, bs, dtype=np.int32)
        lens[-1] = n - (nb - 1) * bs
        off = np.arange(nb, dtype=np.int64) * bs
        bound = N.lib().k4lz4_max_output_size(bs)
        caps = np.full(nb, bound, dtype=np.int32)
        doff = np.arange(nb, dtype=np.int64) * bound
        dst = np.zeros(nb * bound + 16, dtype=np.uint8)
        enc = encode_batch_flat_host(src, off, lens, dst, doff, caps, level, device)
        if (enc <= 0).any():
            raise RuntimeError("Failed to encode chunk. Target buffer too small.")   # LZ4EncoderBase.cs:75-77
        raw = enc >= lens                                       # allowCopy: stored as is, :79-83
        store_len = np.where(raw, lens, enc).astype(np.int32)
        if block_checksum:                                      # checksum of the bytes as stored, :169-175
            cbase = np.concatenate([dst, src]) if raw.any() else dst
            coff = np.where(raw, off + dst.size, doff)
            sums = xxh32_batch(cbase, coff, store_len, 0, device)
        for i in range(nb):
            body = src[off[i]:off[i] + lens[i]] if raw[i] else dst[doff[i]:doff[i] + enc[i]]
            out.append(struct.pack("<I", int(store_len[i]) | (0x80000000 if raw[i] else 0)))   # :159-160
            out.append(body.tobytes())
            if block_checksum:
                out.append(struct.pack("<I", int(sums[i])))
    out.append(struct.pack("<I", 0))                            # end mark, blocking.cs:94
    if content_checksum:
        out.append(struct.pack("<I", xxh32(src)))               # :95
    return b"".join(out)


def read_frame(frame, device: int = 0) -> bytes:
    """Decodes one frame of independent blocks (LZ4FrameReader.blocking.cs:57-144); raises
    InvalidDataException on a bad magic number, header checksum, block or content checksum."""
    f = bytes(frame)
    if len(f) < 7 or struct.unpack_from("<I", f, 0)[0] != MAGIC:
        raise InvalidDataException("LZ4 frame magic number expected")
    flg, bd = f[4], f[5]
    if (flg >> 6) & 0x11 != 1:                                   # sic: the reference masks with 0x11, :85
        raise InvalidDataException(f"LZ4 frame version unknown: {(flg >> 6) & 0x11}")
    chaining = ((flg >> 5) & 1) == 0
    block_checksum = bool((flg >> 4) & 1)
    has_size = bool((flg >> 3) & 1)
    content_checksum = bool((flg >> 2) & 1)
    if flg & 1:
        raise NotImplementedError("Predefined dictionaries feature is not implemented")   # :108-110
    p = 6 + (8 if has_size else 0)
    if len(f) < p + 1 or ((xxh32(f[4:p]) >> 8) & 0xFF) != f[p]:
        raise InvalidDataException("Invalid LZ4 frame header checksum")
    p += 1
    if chaining:
        raise NotImplementedError("chained blocks are decoded by the managed engine (dependent blocks)")
    max_block = _BLOCK_SIZES.get((bd >> 4) & 7, 1 << 16)        # LZ4FrameReader.cs:55-59
    pos, lens, raws, sums = [], [], [], []
    while True:
        if p + 4 > len(f):
            raise InvalidDataException("Unexpected end of stream")
        code = struct.unpack_from("<I", f, p)[0]
        p += 4
        if code == 0:
            break
        blen = code & 0x7FFFFFFF
        if p + blen + (4 if block_checksum else 0) > len(f):
            raise InvalidDataException("Unexpected end of stream")
        pos.append(p); lens.append(blen); raws.append(bool(code & 0x80000000))
        p += blen
        if block_checksum:
            sums.append(struct.unpack_from("<I", f, p)[0])
            p += 4
    expect_content = None
    if content_checksum:
        if p + 4 > len(f):
            raise InvalidDataException("Unexpected end of stream")
        expect_content = struct.unpack_from("<I", f, p)[0]
    nb = len(pos)
    if nb == 0:
        content = b""
    else:
        base = np.frombuffer(f, dtype=np.uint8)
        off = np.array(pos, dtype=np.int64)
        ln = np.array(lens, dtype=np.int32)
        if block_checksum:
            got = xxh32_batch(base, off, ln, 0, device)
            if (got != np.array(sums, dtype=np.uint32)).any():
                raise InvalidDataException("Invalid block checksum")
        is_raw = np.array(raws, dtype=bool)
        caps = np.full(nb, max_block + 8, dtype=np.int32)       # LZ4BlockDecoder.cs:26
        doff = np.arange(nb, dtype=np.int64) * (max_block + 8)
        dst = np.zeros(nb * (max_block + 8) + 16, dtype=np.uint8)
        dec_len = np.where(is_raw, 0, ln).astype(np.int32)       # raw blocks are injected, not decoded
        out_len = decode_batch_flat_host(base, off, dec_len, dst, doff, caps, device)
        parts = []
        for i in range(nb):
            if is_raw[i]:
                if lens[i] > max_block + 8:
                    raise InvalidDataException("block larger than the declared block size")
                parts.append(f[pos[i]:pos[i] + lens[i]])
            else:
                r = int(out_len[i])
                if r < 0 or (r == 0 and lens[i] > 0):
                    raise InvalidDataException("corrupted block")   # InvalidOperationException in LZ4BlockDecoder.cs:50-51
                parts.append(dst[doff[i]:doff[i] + r].tobytes())
        content = b"".join(parts)
    if expect_content is not None and xxh32(content) != expect_content:
        raise InvalidDataException("Invalid content checksum")
    return content
