"""Deterministic test inputs shared by tests/ and tests/golden/make_golden.py.

Modelled on the reference's own test inputs (src/K4os.Compression.LZ4.Tests/
BlockRoundtripTests.cs:63-112: single bytes, repeated bytes incl. 64 KiB of 0xAA, repeated
Lorem text, seeded random bytes) plus sizes around every threshold of the engine
 and blocks of the library's own synthetic generator.
Everything is reproducible from (kind, size, seed) with numpy only.
"""
from __future__ import annotations

import numpy as np

LOREM = (
    b"Sed ut perspiciatis unde omnis iste natus error sit voluptatem accusantium doloremque "
    b"laudantium, totam rem aperiam, eaque ipsa quae ab illo inventore veritatis et quasi "
    b"architecto beatae vitae dicta sunt explicabo. Nemo enim ipsam voluptatem quia voluptas "
    b"sit aspernatur aut odit aut fugit, sed quia consequuntur magni dolores eos qui ratione "
    b"voluptatem sequi nesciunt. Neque porro quisquam est, qui dolorem ipsum quia dolor sit amet. "
)

THRESHOLD_SIZES = [1, 2, 3, 4, 5, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 31, 32, 33, 63, 64, 65,
                   100, 254, 255, 256, 270, 271, 1005, 1006, 1023, 1024, 1025, 4095, 4096, 4097,
                   32767, 65535, 65536]
BIG_SIZES = [65546, 65547, 65548, 100000, 0x123456 // 8]


def gen(kind: str, size: int, seed: int = 0) -> bytes:
    """kind in: random, repeat, lorem, text2, lowent, synth545, synth435, runs."""
    if size == 0:
        return b""
    if kind == "random":
        return np.random.default_rng(seed).integers(0, 256, size, dtype=np.uint8).tobytes()
    if kind == "repeat":
        return bytes([seed & 0xFF]) * size
    if kind == "lorem":
        return (LOREM * (size // len(LOREM) + 1))[:size]
    if kind == "lowent":      # 2-bit symbols: dense accidental matches, many overlapping copies
        return np.random.default_rng(seed).integers(0, 4, size, dtype=np.uint8).tobytes()
    if kind == "text2":       # word soup: realistic short matches at short distances
        rng = np.random.default_rng(seed)
        words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(200)]
        out = bytearray()
        while len(out) < size:
            out += words[int(rng.integers(0, 200))] + b" "
        return bytes(out[:size])
    if kind == "runs":        # long runs of one byte separated by noise: long LSIC lengths, offset 1
        rng = np.random.default_rng(seed)
        out = bytearray()
        while len(out) < size:
            out += bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 3000))
            out += rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8).tobytes()
        return bytes(out[:size])
    if kind.startswith("synth"):
        from k4os.compression.lz4_b200.batch import synth_host
        mp = int(kind[5:])
        nb = (size + 65535) // 65536
        return synth_host(nb, 65536, mp, seed=1234, first_block=seed).tobytes()[:size]
    raise ValueError(kind)


KINDS = ["random", "repeat", "lorem", "lowent", "text2", "runs", "synth525", "synth435"]


def corpus(sizes=None, kinds=None):
    """Yields (name, bytes) over the cross product, small enough to run in seconds."""
    sizes = THRESHOLD_SIZES if sizes is None else sizes
    kinds = KINDS if kinds is None else kinds
    for k in kinds:
        for n in sizes:
            yield f"{k}-{n}", gen(k, n, seed=(n * 7 + 1) & 0xFFFF if k != "repeat" else 0xAA)


def mutate(stream: bytes, rng: np.random.Generator) -> bytes:
    """One random corruption of a compressed stream (for malformed-input parity)."""
    c = bytearray(stream)
    k = int(rng.integers(0, 5))
    if k == 0 and len(c):
        for _ in range(int(rng.integers(1, 4))):
            c[int(rng.integers(0, len(c)))] = int(rng.integers(0, 256))
    elif k == 1 and len(c) > 1:
        c = c[:int(rng.integers(1, len(c)))]
    elif k == 2:
        c += rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8).tobytes()
    elif k == 3 and len(c):
        c[int(rng.integers(0, len(c)))] = [0xFF, 0xF0, 0x0F, 0x00][int(rng.integers(0, 4))]
    elif len(c) > 4:
        i = int(rng.integers(0, len(c) - 2))
        c[i + 1] = 0; c[i + 2] = 0          # plant a zero offset somewhere
    return bytes(c)


def uses_zero_offset(stream: bytes) -> bool:
    """True iff the token chain of `stream` contains a match with offset 0 (or cannot be walked).

    The reference accepts such a match -- `match == op >= dst` passes LL64.dec.cs:338 -- and copies
    whatever the destination buffer held before, so the CONTENT of the decoded block is unspecified
    (only the returned length is contractual).  Content comparisons skip exactly these streams; a
    plain `b"\x00\x00" in stream` test would also skip every stream with two zero literals."""
    c = bytes(stream)
    n = len(c)
    p = 0
    try:
        while p < n:
            t = c[p]; p += 1
            lit = t >> 4
            if lit == 15:
                while True:
                    s = c[p]; p += 1; lit += s
                    if s != 255:
                        break
            p += lit
            if p + 2 > n:
                return False
            if c[p] == 0 and c[p + 1] == 0:
                return True
            p += 2
            if (t & 15) == 15:
                while True:
                    s = c[p]; p += 1
                    if s != 255:
                        break
        return False
    except IndexError:
        return True
