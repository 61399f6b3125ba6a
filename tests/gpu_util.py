"""Helpers for the -m gpu tests: move batches through the device-resident C ABI."""
import numpy as np
import torch

from pg_cryogen_b200.codec import CRYO_BLCKSZ, pack_chunks


def decode_device(gpu, methods, chunks, block_size=CRYO_BLCKSZ, fill=0x5A):
    """Decompress a batch through cryogpu_decompress_device.
    Returns (out [n, block_size] uint8, out_size, status) as numpy arrays."""
    n = len(chunks)
    buf, offs, sizes = pack_chunks(chunks)
    dev = torch.device("cuda", gpu.device)
    methods = np.ascontiguousarray(np.broadcast_to(np.asarray(methods, dtype=np.int32), (n,)))
    d_src = torch.from_numpy(buf).to(dev)
    d_off = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_sz = torch.from_numpy(sizes.view(np.int32)).to(dev)
    d_me = torch.from_numpy(methods).to(dev)
    d_dst = torch.full((n, block_size), fill, dtype=torch.uint8, device=dev)
    d_osz = torch.full((n,), -1, dtype=torch.int32, device=dev)
    d_st = torch.full((n,), -1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    gpu.decompress_device(d_me, d_src, d_off, d_sz, d_dst, block_size, d_osz, d_st, n,
                          block_size=block_size, stream=stream)
    torch.cuda.synchronize(dev)
    return d_dst.cpu().numpy(), d_osz.cpu().numpy().view(np.uint32), d_st.cpu().numpy()


def encode_device(gpu, method, level, blocks, block_size=CRYO_BLCKSZ):
    """Compress [n, block_size] blocks through cryogpu_compress_device.
    Returns (list of compressed uint8 arrays, status)."""
    from pg_cryogen_b200.codec import compress_bound
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, block_size)
    n = blocks.shape[0]
    dev = torch.device("cuda", gpu.device)
    bound = compress_bound(method, block_size)
    stride = (bound + 15) & ~15
    d_src = torch.from_numpy(blocks).to(dev)
    d_dst = torch.zeros((n, stride), dtype=torch.uint8, device=dev)
    d_sz = torch.full((n,), -1, dtype=torch.int32, device=dev)
    d_st = torch.full((n,), -1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    gpu.compress_device(method, level, d_src, block_size, d_dst, stride, stride, d_sz, d_st, n,
                        block_size=block_size, stream=stream)
    torch.cuda.synchronize(dev)
    sz = d_sz.cpu().numpy().view(np.uint32)
    out = d_dst.cpu().numpy()
    return [out[i, : sz[i]].copy() for i in range(n)], d_st.cpu().numpy()
