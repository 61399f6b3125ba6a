"""-m gpu: cryogpu_decompress_pages_* / cryogpu_compress_pages_* against the restatement of the reference's
page split and gather (oracle/cryo_pages.c: pg_cryogen.c:689-805, cache.c:100-176) and the reference's
codec (oracle/_ref).  Chains of 1, 2 and 129 pages, adjacent and scattered over the relation; the page
errors of cache.c:112-129; page images byte-for-byte."""
import numpy as np
import pytest
import torch

from oracle import pages as opg
from pg_cryogen_b200 import CRYO_BLCKSZ, blockgen as bg
from pg_cryogen_b200.codec import compress_bound, pages_needed

pytestmark = pytest.mark.gpu
PAGE = 8192


def _relation(oracle_ref, rng, scattered):
    """Cryo blocks of every kind written into a relation by the reference's split.  -> (rel, blocks, comp, methods, chains)"""
    blocks = [bg.make_block("S", "hex", 11), bg.make_block("S", "lowcard", 12), bg.make_block("M", "hex", 13),
              rng.integers(0, 256, size=CRYO_BLCKSZ, dtype=np.uint8), bg.make_block("D", "lowcard", 15), np.zeros(CRYO_BLCKSZ, dtype=np.uint8),
              bg.regression_block(1, 290)]
    methods = [0, 0, 1, 0, 1, 1, 0]
    comp = [oracle_ref.compress(m, 1, b)[0][0] for m, b in zip(methods, blocks)]
    total = sum(opg.pages_needed(c.size) for c in comp)
    rel = np.zeros((total + 40, PAGE), dtype=np.uint8)
    free = list(rng.permutation(np.arange(1, rel.shape[0]))) if scattered else list(range(rel.shape[0] - 1, 0, -1))
    chains = []
    for i, c in enumerate(comp):
        ch = [int(free.pop()) for _ in range(opg.pages_needed(c.size))]
        opg.split(rel, ch, c, methods[i], 700 + i)
        chains.append(ch)
    return rel, blocks, comp, methods, chains, free


@pytest.mark.parametrize("scattered", [False, True])
def test_decompress_from_page_chains(gpu, oracle_ref, scattered):
    rng = np.random.default_rng(5)
    rel, blocks, comp, methods, chains, free = _relation(oracle_ref, rng, scattered)
    lens = sorted(len(ch) for ch in chains)
    assert lens[0] == 1 and 2 in lens and lens[-1] >= 129          # one page, two pages, an incompressible block
    out, osz, st, me, csz = gpu.decompress_pages_host(rel, chains)
    assert (st == 0).all(), st
    assert list(me) == methods and list(csz) == [c.size for c in comp]
    for i, b in enumerate(blocks):
        assert osz[i] == CRYO_BLCKSZ and np.array_equal(out[i], b), i


def test_page_chain_errors_are_per_block(gpu, oracle_ref):
    rng = np.random.default_rng(6)
    rel, blocks, comp, methods, chains, free = _relation(oracle_ref, rng, True)
    bad = [list(ch) for ch in chains]
    rel = rel.copy()
    bad[0] = [int(free[0])]                     # a new page (pd_upper == 0): CRYO_ERR_EMPTY_BLOCK, cache.c:115
    bad[2] = chains[2][1:]                      # starts at the second page: CRYO_ERR_WRONG_STARTING_BLOCK, cache.c:125
    bad[3] = chains[3][:50]                     # the reader's walk ended early: too few pages for compressed_size
    rel[chains[4][2], 28:32] = 0xFF             # a page whose `next` does not name the page the host read after it
    bad[6] = []                                 # no pages at all
    out, osz, st, me, csz = gpu.decompress_pages_host(rel, bad)
    assert st[0] == 8 and st[2] == 9 and st[3] == 10 and st[4] == 10 and st[6] == 8, st
    assert opg.gather(rel, bad[0][0])[0] == opg.ERR_EMPTY_BLOCK
    assert opg.gather(rel, bad[2][0])[0] == opg.ERR_WRONG_STARTING_BLOCK
    for i in (1, 5):
        assert st[i] == 0 and np.array_equal(out[i], blocks[i]), i
    for i in (0, 2, 3, 4, 6):
        assert osz[i] == 0


@pytest.mark.parametrize("method,level", [(0, 1), (1, 1), (1, -3)])
def test_compress_into_page_chains(gpu, oracle_ref, method, level):
    """The GPU's page images are what the reference's split makes of the GPU's compressed bytes, and the
    reference reads the blocks back through its own gather + cryo_decompress."""
    blocks = np.stack([bg.make_block("S", "hex", 21), bg.make_block("M", "lowcard", 22), bg.make_block("D", "random", 23),
                       np.zeros(CRYO_BLCKSZ, dtype=np.uint8), bg.make_block("D", "hex", 24)])
    n = blocks.shape[0]
    cap = pages_needed(compress_bound(method))
    rng = np.random.default_rng(9)
    rel = np.full((n * cap + 8, PAGE), 0xA5, dtype=np.uint8)        # stale bytes: every page must be written whole
    blknos = rng.permutation(np.arange(1, rel.shape[0]))[: n * cap].astype(np.uint32).reshape(n, cap)
    npg, csz, st = gpu.compress_pages_host(method, level, blocks, rel, blknos, created_xid=4242)
    assert (st == 0).all(), st
    want_rel = np.zeros_like(rel)
    for i in range(n):
        assert npg[i] == opg.pages_needed(int(csz[i]))
        err, m, size, got, chain = opg.gather(rel, int(blknos[i, 0]))
        assert err == opg.ERR_SUCCESS and m == method and size == csz[i] and chain == [int(x) for x in blknos[i, : npg[i]]]
        back, ok = oracle_ref.decompress_one(method, got)
        assert ok and np.array_equal(back, blocks[i]), i
        opg.split(want_rel, blknos[i, : npg[i]], got, method, 4242)
        for k in range(int(npg[i])):
            b = int(blknos[i, k])
            assert np.array_equal(rel[b], want_rel[b]), (i, k)
        for k in range(int(npg[i]), cap):
            assert (rel[int(blknos[i, k])] == 0xA5).all()           # pages the block did not need are untouched


def test_pages_device_roundtrip_many_blocks(gpu, oracle_ref):
    """Device-resident: compress 64 blocks into pages, decompress from those pages (slots permuted)."""
    dev = torch.device("cuda", gpu.device)
    uniq = np.stack([bg.make_block(k, p, 30 + i) for i, (k, p) in enumerate((("S", "hex"), ("M", "hex"), ("D", "lowcard"), ("S", "lowcard")))])
    n = 64
    d_src = torch.from_numpy(uniq).to(dev).repeat(n // 4, 1).contiguous()
    for method in (0, 1):
        cap = pages_needed(compress_bound(method))
        blk = torch.arange(1, n * cap + 1, dtype=torch.int32, device=dev)
        d_pages = torch.zeros((n * cap, PAGE), dtype=torch.uint8, device=dev)
        d_np = torch.zeros(n, dtype=torch.int32, device=dev)
        d_cs = torch.zeros(n, dtype=torch.int32, device=dev)
        d_st = torch.full((n,), -1, dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        L = gpu.lib
        rc = L.cryogpu_compress_pages_device(gpu.handle, n, method, 1, d_src.data_ptr(), CRYO_BLCKSZ, CRYO_BLCKSZ, blk.data_ptr(),
                                             cap, 77, d_pages.data_ptr(), d_np.data_ptr(), d_cs.data_ptr(), d_st.data_ptr(), s or 1)
        assert rc == 0 and bool((d_st == 0).all())
        npg = d_np.cpu().numpy()
        # chains: block i's pages are slots i * cap + k, block numbers 1 + i * cap + k
        slots = np.concatenate([np.arange(i * cap, i * cap + npg[i]) for i in range(n)]).astype(np.uint32)
        coff = np.zeros(n + 1, dtype=np.uint32)
        coff[1:] = np.cumsum(npg)
        d_slot = torch.from_numpy(slots.view(np.int32)).to(dev)
        d_blk = torch.from_numpy((slots + 1).view(np.int32)).to(dev)
        d_coff = torch.from_numpy(coff.view(np.int32)).to(dev)
        d_out = torch.zeros((n, CRYO_BLCKSZ), dtype=torch.uint8, device=dev)
        d_osz = torch.zeros(n, dtype=torch.int32, device=dev)
        d_me = torch.full((n,), -1, dtype=torch.int32, device=dev)
        d_st.fill_(-1)
        rc = L.cryogpu_decompress_pages_device(gpu.handle, n, d_pages.data_ptr(), d_slot.data_ptr(), d_blk.data_ptr(), d_coff.data_ptr(),
                                               int(slots.size), d_out.data_ptr(), CRYO_BLCKSZ, CRYO_BLCKSZ, d_osz.data_ptr(),
                                               d_st.data_ptr(), d_me.data_ptr(), None, s or 1)
        assert rc == 0
        torch.cuda.synchronize(dev)
        assert bool((d_st == 0).all()) and bool((d_me == method).all()) and bool((d_osz == CRYO_BLCKSZ).all())
        assert bool(torch.equal(d_out, d_src))
