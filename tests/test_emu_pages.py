"""The page-chain and tuple-walk kernel bodies (cryo_pages.cuh) under the SIMT emulator (tests/emu/emu_pages.cpp),
launched as cryogpu.cu launches them, against the restatement of the reference's split (pg_cryogen.c:689-805),
gather (cache.c:100-176) and item walk (pg_cryogen.c:293, storage.c:55-68) in oracle/cryo_pages.c.  CPU only; the
same comparisons run on the device in tests/test_gpu_pages.py and tests/test_gpu_tuples.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pages as opg
from pg_cryogen_b200 import blockgen as bg

HERE = os.path.dirname(os.path.abspath(__file__))
PAGE = 8192
ST_EMPTY, ST_WRONG_START, ST_CHAIN, METHOD_SKIP = 8, 9, 10, 0x7FFFFFFF


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    return C.CDLL(os.path.join(HERE, "emu", "libcryoemu_pages.so"))


def _aligned(n, fill=0):
    raw = np.full(n + 64, fill, dtype=np.uint8)
    at = (-raw.ctypes.data) % 16
    return raw[at: at + n]


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _gather(emu, rel, chains, max_csize=(1 << 20) + PAGE):
    """chains: per cryo block the block numbers the host walked; the relation is the page buffer (slot = block number)"""
    n = len(chains)
    flat = np.array([b for ch in chains for b in ch] + [0], dtype=np.uint32)
    off = np.zeros(n + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(ch) for ch in chains])
    pages = _aligned(rel.size)
    pages[:] = rel.reshape(-1)
    comp = _aligned((int(off[-1]) + 1) * PAGE, 0xAA)
    src_off, src_size = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint32)
    dec, hdr, st = (np.zeros(n, dtype=np.int32) for _ in range(3))
    rc = emu.emu_pages_gather(_ptr(pages), _ptr(flat), _ptr(flat), _ptr(off), n, _ptr(comp), _ptr(src_off), _ptr(src_size),
                              _ptr(dec), _ptr(hdr), _ptr(st), C.c_uint32(max_csize))
    assert rc == 0
    return [(int(st[b]), int(dec[b]), int(src_size[b]), comp[int(src_off[b]): int(src_off[b]) + int(src_size[b])].copy())
            for b in range(n)]


def test_emulated_gather_equals_the_restatement(emu, oracle_ref):
    rng = np.random.default_rng(11)
    rel = np.zeros((700, PAGE), dtype=np.uint8)
    free = list(rng.permutation(np.arange(1, 700)))
    streams = [oracle_ref.compress(i & 1, 1, bg.make_block(k, p, i))[0][0]
               for i, (k, p) in enumerate((("S", "hex"), ("M", "hex"), ("D", "random"), ("M", "lowcard")))]
    streams += [rng.integers(0, 256, size=s, dtype=np.uint8) for s in (1, 15, 8144, 8145, 8144 + 8160, 8144 + 8160 + 1, 40_000)]
    chains = []
    for i, c in enumerate(streams):
        ch = [int(free.pop()) for _ in range(opg.pages_needed(c.size))]
        assert opg.split(rel, ch, c, i & 1, 500 + i) == len(ch)
        chains.append(ch)
    got = _gather(emu, rel, chains)
    for i, c in enumerate(streams):
        err, method, size, want, chain = opg.gather(rel, chains[i][0])
        assert err == opg.ERR_SUCCESS and chain == chains[i] and np.array_equal(want, c)
        assert got[i][:3] == (0, method, size) and np.array_equal(got[i][3], want), i
    # a chain longer than the block needs (the host walked on into another block's pages): the extra entries are ignored
    longer = _gather(emu, rel, [chains[2] + chains[0][:2]])[0]
    assert longer[:3] == (0, 0, streams[2].size) and np.array_equal(longer[3], streams[2])
    # cache.c:115-129: a page never written; a chain entered in its middle
    never = int(free[0])
    bad = _gather(emu, rel, [[never], chains[1][1:], [], chains[3]])
    assert opg.gather(rel, never)[0] == opg.ERR_EMPTY_BLOCK and bad[0][:3] == (ST_EMPTY, METHOD_SKIP, 0)
    assert opg.gather(rel, chains[1][1])[0] == opg.ERR_WRONG_STARTING_BLOCK and bad[1][:3] == (ST_WRONG_START, METHOD_SKIP, 0)
    assert bad[2][:3] == (ST_EMPTY, METHOD_SKIP, 0)
    assert bad[3][0] == 0 and np.array_equal(bad[3][3], streams[3])         # the neighbours of failed chains are served
    # the chain is cut (next = invalid on its fourth page): the restatement comes back short and the reference then fails in
    # cryo_decompress (cache.c:163-178); the host's walk ends there too, and the gather reports the chain
    rel2 = rel.copy()
    rel2[chains[2][3], 28:32] = 0xFF
    err, method, size, short, walked = opg.gather(rel2, chains[2][0])
    assert err == opg.ERR_SUCCESS and short.size < size and walked == chains[2][:4]
    assert _gather(emu, rel2, [walked])[0][:3] == (ST_CHAIN, METHOD_SKIP, 0)
    # a page of the chain names another first page / the next page is not the one the host handed over
    rel3 = rel.copy()
    rel3[chains[0][1], 24:28] = 7
    assert _gather(emu, rel3, [chains[0]])[0][0] == ST_CHAIN
    swapped = chains[0][:1] + chains[0][2:3] + chains[0][1:2] + chains[0][3:]
    assert _gather(emu, rel, [swapped])[0][0] == ST_CHAIN
    # compressed_size beyond what a block can be
    assert _gather(emu, rel, [chains[2]], max_csize=streams[2].size - 1)[0][0] == ST_CHAIN


@pytest.mark.parametrize("method, xid", [(0, 1), (1, 0xFFFFFFF0)])
def test_emulated_split_writes_the_restatements_pages(emu, method, xid):
    rng = np.random.default_rng(12 + method)
    sizes = [1, 2, 15, 16, 17, 8143, 8144, 8145, 8144 + 8160 - 1, 8144 + 8160, 8144 + 8160 + 1, 100_000, 3 * 8160 + 8144 + 5]
    n, cap_pages, stride = len(sizes), 14, 100_016
    comp = _aligned(n * stride, 0x5A)
    for b, s in enumerate(sizes):
        comp[b * stride: b * stride + s] = rng.integers(0, 256, size=s, dtype=np.uint8)
    blkno = rng.permutation(np.arange(1, 1 + n * cap_pages)).astype(np.uint32)
    out = _aligned(n * cap_pages * PAGE, 0xCC)
    npages = np.zeros(n, dtype=np.uint32)
    rc = emu.emu_pages_split(_ptr(comp), C.c_uint64(stride), _ptr(np.array(sizes, dtype=np.uint32)), n, method, C.c_uint32(xid),
                             _ptr(blkno), cap_pages, _ptr(out), C.c_uint64(cap_pages * PAGE), _ptr(npages))
    assert rc == 0
    rel = np.zeros((2 + n * cap_pages, PAGE), dtype=np.uint8)
    for b, s in enumerate(sizes):
        need = opg.pages_needed(s)
        ch = blkno[b * cap_pages: (b + 1) * cap_pages]
        if need > cap_pages:
            assert npages[b] == 0, s
            continue
        assert opg.split(rel, ch[:need], comp[b * stride: b * stride + s], method, xid) == need == npages[b]
        img = out[b * cap_pages * PAGE: (b * cap_pages + need) * PAGE].reshape(need, PAGE)
        for k in range(need):
            assert np.array_equal(img[k], rel[ch[k]]), (s, k, int(np.argmax(img[k] != rel[ch[k]])))
        # nothing beyond the pages it needs
        assert np.all(out[(b * cap_pages + need) * PAGE: (b + 1) * cap_pages * PAGE] == 0xCC)


def test_emulated_tuple_walk_equals_the_restatement(emu):
    blocks = [bg.make_block(k, p, s) for k, p, s in (("S", "hex", 1), ("M", "lowcard", 2), ("D", "hex", 3), ("D", "random", 4),
                                                     ("S", "lowcard", 5), ("M", "random", 6), ("D", "lowcard", 7),
                                                     ("S", "random", 8), ("M", "hex", 9))]
    bad = blocks[0].copy()
    bad[8 + 8 * 7: 8 + 8 * 7 + 4] = 0xFF                    # item 8 points outside the block
    zeros = np.zeros(1 << 20, dtype=np.uint8)               # lower = 0: no items, not a block the walk accepts
    huge = blocks[1].copy()
    huge[0:4] = 0xFF                                        # lower beyond the block
    blocks += [bad, zeros, huge]
    n = len(blocks)
    buf = _aligned(n << 20)
    for b, blk in enumerate(blocks):
        buf[b << 20: (b + 1) << 20] = blk
    nt, by, ok = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.int32)
    assert emu.emu_tuple_stats(_ptr(buf), C.c_uint64(1 << 20), 1 << 20, n, _ptr(nt), _ptr(by), _ptr(ok)) == 0
    for b, blk in enumerate(blocks):
        assert (int(nt[b]), int(by[b]), int(ok[b])) == opg.block_tuple_stats(blk), b
    assert ok[9] == 0 and ok[0] == 1
