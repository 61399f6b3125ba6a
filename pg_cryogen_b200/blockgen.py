"""Synthetic cryo-block generator (the plaintext the codec sees).

A cryo block is CRYO_BLCKSZ = 1 MiB (reference storage.h:18) laid out by the
reference's storage.c:15-50 (``cryo_init_page`` / ``cryo_storage_insert``):

    offset 0      uint32 lower = 8 + 8*ntuples
    offset 4      uint32 upper = offset of the lowest tuple
    offset 8      CryoItemId[ntuples] {uint32 off; uint32 len}    (storage.h:73-77)
    lower..upper  zeros
    upper..1Mi    tuples packed downward, each start MAXALIGN(8)-aligned

with at most 290 tuples per block (storage.c:10, :32-33).  This module restates
that packing in numpy so the benchmark and the GPU tests do not need the
reference at run time; tests/test_blockgen.py checks it byte-for-byte against
the reference's own storage.c (through oracle/_ref).

Tuple images follow SURVEY.md section 8(d): a 24-byte heap tuple header as
``heap_form_tuple`` leaves it, then the attribute data.  Three block kinds
(S sparse, M medium, D dense) and three payload distributions (hex, lowcard,
random).  All randomness is a counter-based splitmix64 keyed on
``seed + block_index`` so that any block can be generated independently on any
rank.
"""
from __future__ import annotations

import hashlib

import numpy as np

CRYO_BLCKSZ = 1 << 20
MAX_TUPLES = 290            # storage.c:32-33 with MaxHeapTuplesPerPage = 291
DATA_HEADER = 8             # CryoDataHeaderSize, storage.h:86
ITEMID = 8                  # sizeof(CryoItemId), storage.h:73-77
SEED = 0x9E3779B97F4A7C15

# kind -> (tuples per block, t_len)
KINDS = {
    "S": (290, 61),         # the regression table's row shape (sql/pg_cryogen.sql:3-8)
    "M": (290, 1124),       # jsonb-like rows (sql/pg_cryogen.sql:62-93)
    "D": (288, 3624),       # 99.8 % full
}
PAYLOADS = ("hex", "lowcard", "random")

_HEX = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)
_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def maxalign(n: int) -> int:
    return (n + 7) & ~7


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser over uint64 counters."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        return z ^ (z >> np.uint64(31))


def rand_u64(seed: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) * np.uint64(0x2545F4914F6CDD1D)
        ctr = np.arange(n, dtype=np.uint64) + base
    return splitmix64(ctr)


def rand_bytes(seed: int, n: int) -> np.ndarray:
    return rand_u64(seed, (n + 7) // 8).view(np.uint8)[:n]


def _word_dict() -> list[bytes]:
    r = rand_u64(0xD1C7, 200 * 12).reshape(200, 12)
    words = []
    for i in range(200):
        ln = 3 + int(r[i, 0] % np.uint64(8))
        words.append(bytes(int(97 + (r[i, 1 + j] % np.uint64(26))) for j in range(ln)))
    return words


_WORDS = _word_dict()
_WORD_ARR = [np.frombuffer(w + b" ", dtype=np.uint8) for w in _WORDS]


def heap_header(t_len: int, natts: int, typeid: int = 0x4000) -> np.ndarray:
    """The 24-byte HeapTupleHeader image heap_form_tuple produces."""
    h = np.zeros(24, dtype=np.uint8)
    h[0:4] = np.frombuffer(np.uint32(t_len << 2).tobytes(), dtype=np.uint8)   # datum_len_
    h[4:8] = 0xFF                                                             # typmod -1
    h[8:12] = np.frombuffer(np.uint32(typeid).tobytes(), dtype=np.uint8)      # typeid
    h[12:16] = 0xFF                                                           # invalid ctid
    h[18:20] = np.frombuffer(np.uint16(natts).tobytes(), dtype=np.uint8)      # t_infomask2
    h[20:22] = np.frombuffer(np.uint16(0x0002).tobytes(), dtype=np.uint8)     # HEAP_HASVARWIDTH
    h[22] = 24                                                                # t_hoff
    return h


def _text_fill(dst: np.ndarray, payload: str, seed: int) -> None:
    """Fill the 2-D uint8 array dst[n, w] with text of the given distribution."""
    n, w = dst.shape
    if w == 0:
        return
    if payload == "hex":
        nib = rand_bytes(seed, (n * w + 1) // 2)
        both = np.empty(nib.size * 2, dtype=np.uint8)
        both[0::2] = nib & 0x0F
        both[1::2] = nib >> 4
        dst[:] = _HEX[both[: n * w]].reshape(n, w)
    elif payload == "random":
        dst[:] = rand_bytes(seed, n * w).reshape(n, w)
    elif payload == "lowcard":
        # words from a 200-entry dictionary; enough picks to cover the widest row
        picks_per_row = w // 4 + 2
        idx = (rand_u64(seed, n * picks_per_row) % np.uint64(200)).astype(np.int64)
        idx = idx.reshape(n, picks_per_row)
        for r in range(n):
            row = np.concatenate([_WORD_ARR[i] for i in idx[r]])
            dst[r] = row[:w]
    else:
        raise ValueError(f"unknown payload {payload!r}")


def make_tuples(kind: str, payload: str, block_index: int, ntuples: int | None = None,
                seed: int = SEED) -> np.ndarray:
    """Return the tuple images of one block as a uint8 matrix [ntuples, t_len]."""
    n_full, t_len = KINDS[kind]
    n = n_full if ntuples is None else ntuples
    bseed = (seed + block_index) & 0xFFFFFFFFFFFFFFFF
    ids = (np.arange(n, dtype=np.uint32) + np.uint32((block_index * n_full + 1) & 0xFFFFFFFF))
    t = np.zeros((n, t_len), dtype=np.uint8)
    lowcard = payload == "lowcard"
    natts = 3 if lowcard else 2
    t[:, :24] = heap_header(t_len, natts)
    t[:, 24:28] = ids.view(np.uint8).reshape(n, 4)
    pos = 28
    if lowcard:
        ts = (np.uint64(1_600_000_000_000_000) + np.uint64(block_index) * np.uint64(n_full * 1000)
              + np.arange(n, dtype=np.uint64) * np.uint64(1000))
        t[:, 32:40] = ts.view(np.uint8).reshape(n, 8)
        pos = 40
    remaining = t_len - pos
    if remaining <= 127:
        # short varlena: 1-byte header (len << 1) | 1, len includes the header
        t[:, pos] = (remaining << 1) | 1
        _text_fill(t[:, pos + 1:], payload, bseed)
    else:
        # 4-byte varlena header: len << 2 (little-endian), len includes the header
        t[:, pos:pos + 4] = np.frombuffer(np.uint32(remaining << 2).tobytes(), dtype=np.uint8)
        _text_fill(t[:, pos + 4:], payload, bseed)
    return t


def pack_block(tuples, out: np.ndarray | None = None) -> np.ndarray:
    """Pack tuple images into one cryo block exactly as storage.c:15-50 does.

    ``tuples`` is a uint8 matrix [n, t_len] or a sequence of bytes objects.
    Tuples that no longer fit (storage.c:32-37) raise ValueError.
    """
    blk = np.zeros(CRYO_BLCKSZ, dtype=np.uint8) if out is None else out
    if out is not None:
        blk[:] = 0
    if isinstance(tuples, np.ndarray) and tuples.ndim == 2:
        n, t_len = tuples.shape
        step = maxalign(t_len)
        if n > MAX_TUPLES:
            raise ValueError("more than 290 tuples per block")
        if n and (t_len + ITEMID) > (CRYO_BLCKSZ - (n - 1) * step) - (DATA_HEADER + (n - 1) * ITEMID):
            raise ValueError("tuples do not fit")
        offs = CRYO_BLCKSZ - step * np.arange(1, n + 1, dtype=np.int64)
        if n:
            # tuple k lives at [offs[k], offs[k] + t_len); rows are written highest-first
            area = blk[CRYO_BLCKSZ - n * step:].reshape(n, step)
            area[::-1, :t_len] = tuples
        item = np.empty((n, 2), dtype=np.uint32)
        item[:, 0] = offs
        item[:, 1] = t_len
        blk[DATA_HEADER:DATA_HEADER + n * ITEMID] = item.view(np.uint8).reshape(-1)
        lower = DATA_HEADER + n * ITEMID
        upper = CRYO_BLCKSZ - n * step
    else:
        lower, upper = DATA_HEADER, CRYO_BLCKSZ
        for k, tup in enumerate(tuples):
            tl = len(tup)
            if (tl + ITEMID) > (upper - lower) or ((lower - DATA_HEADER) // ITEMID) + 1 >= MAX_TUPLES + 1:
                raise ValueError(f"tuple {k} does not fit")
            upper -= maxalign(tl)
            blk[upper:upper + tl] = np.frombuffer(bytes(tup), dtype=np.uint8)
            blk[lower:lower + 8] = np.array([upper, tl], dtype=np.uint32).view(np.uint8)
            lower += ITEMID
    blk[0:8] = np.array([lower, upper], dtype=np.uint32).view(np.uint8)
    return blk


def make_block(kind: str, payload: str, block_index: int, ntuples: int | None = None,
               seed: int = SEED, out: np.ndarray | None = None) -> np.ndarray:
    return pack_block(make_tuples(kind, payload, block_index, ntuples, seed), out=out)


def make_blocks(kind: str, payload: str, first_index: int, count: int,
                seed: int = SEED, out: np.ndarray | None = None) -> np.ndarray:
    """Blocks first_index .. first_index+count-1 as a [count, 1 MiB] uint8 array."""
    arr = np.empty((count, CRYO_BLCKSZ), dtype=np.uint8) if out is None else out
    for i in range(count):
        make_block(kind, payload, first_index + i, seed=seed, out=arr[i])
    return arr


def make_table_blocks(nrows: int, kind: str = "S", payload: str = "hex",
                      first_block: int | None = None, count: int | None = None,
                      seed: int = SEED, out: np.ndarray | None = None) -> np.ndarray:
    """Blocks of an ``nrows``-row table (last block partially filled).

    ``first_block``/``count`` select a contiguous block range (multi-GPU shards).
    """
    per = KINDS[kind][0]
    total = (nrows + per - 1) // per
    lo = 0 if first_block is None else first_block
    cnt = total - lo if count is None else count
    arr = np.empty((cnt, CRYO_BLCKSZ), dtype=np.uint8) if out is None else out
    for i in range(cnt):
        b = lo + i
        n = min(per, nrows - b * per)
        make_block(kind, payload, b, ntuples=n, seed=seed, out=arr[i])
    return arr


def table_block_count(nrows: int, kind: str = "S") -> int:
    per = KINDS[kind][0]
    return (nrows + per - 1) // per


def regression_block(first_id: int, last_id: int) -> np.ndarray:
    """Rows ``first_id..last_id`` of the regression table (sql/pg_cryogen.sql:3-9):
    (id int4, md5(id::text) as text) -> the S tuple shape with a real md5."""
    n = last_id - first_id + 1
    t = np.zeros((n, 61), dtype=np.uint8)
    t[:, :24] = heap_header(61, 2)
    for k in range(n):
        i = first_id + k
        t[k, 24:28] = np.frombuffer(np.uint32(i).tobytes(), dtype=np.uint8)
        t[k, 28] = (33 << 1) | 1
        t[k, 29:61] = np.frombuffer(hashlib.md5(str(i).encode()).hexdigest().encode(), dtype=np.uint8)
    return pack_block(t)
