"""dev probe: time batched decode on the GPU (not a bench; see bench.py)."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
sys.path.insert(0, ".")
from pg_cryogen_b200 import CryoGPU, blockgen as bg
from pg_cryogen_b200.codec import pack_chunks
from oracle import ref

_shared = None

def probe(method, level, kind, payload, n):
    global _shared
    if os.environ.get("PROBE_SHARED_CTX"):
        _shared = _shared or CryoGPU(0)
        g = _shared
    else:
        g = CryoGPU(0)
    uniq = min(n, 32)
    blocks = bg.make_blocks(kind, payload, 0, uniq)
    comp, sizes, _ = ref.compress(method, level, blocks, nthreads=8)
    chunks = [comp[i % uniq] for i in range(n)]
    buf, offs, sz = pack_chunks(chunks)
    dev = torch.device("cuda:0")
    d_src = torch.from_numpy(buf).to(dev); d_off = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_sz = torch.from_numpy(sz.view(np.int32)).to(dev)
    d_me = torch.full((n,), method, dtype=torch.int32, device=dev)
    d_dst = torch.zeros((n, 1 << 20), dtype=torch.uint8, device=dev)
    d_osz = torch.zeros((n,), dtype=torch.int32, device=dev); d_st = torch.full((n,), -1, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        g.decompress_device(d_me, d_src, d_off, d_sz, d_dst, 1 << 20, d_osz, d_st, n, stream=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = int(os.environ.get("PROBE_REPS", "5"))
    e0.record()
    for _ in range(reps):
        g.decompress_device(d_me, d_src, d_off, d_sz, d_dst, 1 << 20, d_osz, d_st, n, stream=s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ok = bool((d_st == 0).all().item())
    good = np.array_equal(d_dst[:uniq].cpu().numpy(), blocks)
    gbs = n * (1 << 20) / ms / 1e6
    alg = (n * (1 << 20) + int(sz.astype(np.int64).sum())) / ms / 1e6
    print(f"method={method} level={level} {kind}/{payload} n={n}: {ms:.3f} ms  {gbs:.1f} GB/s out, {alg:.1f} GB/s algorithmic, status_ok={ok} exact={good}" + (f" dst=0x{d_dst.data_ptr():x} src=0x{d_src.data_ptr():x}" if os.environ.get("PROBE_PTRS") else ""), flush=True)

if __name__ == "__main__":
    # usage: gpu_probe.py N [method:level:kind:payload ...]
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    specs = sys.argv[2:] or ["0:1:S:hex", "0:1:M:hex", "0:1:D:hex", "0:1:D:lowcard",
                             "1:1:S:hex", "1:1:M:hex", "1:1:D:hex", "1:1:D:lowcard"]
    for sp in specs:
        m, l, k, pl = sp.split(":")
        try:
            probe(int(m), int(l), k, pl, n)
        except Exception as ex:
            print("probe failed", sp, ex, flush=True)
