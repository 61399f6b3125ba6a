"""dev probe (BASELINE.json configs[4]): latency of small mixed batches through cryogpu_decompress_device,
p50 / p99 over 200 calls per batch size, for the default zstd path and the warp-per-frame kernel."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
sys.path.insert(0, ".")
from pg_cryogen_b200 import CryoGPU, blockgen as bg
from pg_cryogen_b200.codec import pack_chunks
from oracle import ref

def main():
    g = CryoGPU(0)
    dev = torch.device("cuda:0")
    kinds = [("S", "hex"), ("M", "hex"), ("D", "lowcard"), ("S", "lowcard")]
    blocks = np.stack([bg.make_block(k, p, i) for i, (k, p) in enumerate(kinds * 8)])
    z, _, _ = ref.compress(1, 1, blocks, nthreads=8)
    l, _, _ = ref.compress(0, 1, blocks, nthreads=8)
    tag = os.environ.get("CRYOGPU_ZSTD_KERNEL", "pipeline")
    for n in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        chunks, methods = [], []
        for i in range(n):
            m = i & 1
            chunks.append((z if m else l)[i % len(z)])
            methods.append(m)
        buf, offs, sz = pack_chunks(chunks)
        d_src = torch.from_numpy(buf).to(dev); d_off = torch.from_numpy(offs.view(np.int64)).to(dev)
        d_sz = torch.from_numpy(sz.view(np.int32)).to(dev); d_me = torch.tensor(methods, dtype=torch.int32, device=dev)
        d_dst = torch.empty((n, 1 << 20), dtype=torch.uint8, device=dev)
        d_osz = torch.zeros((n,), dtype=torch.int32, device=dev); d_st = torch.full((n,), -1, dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        ts = []
        for it in range(220):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g.decompress_device(d_me, d_src, d_off, d_sz, d_dst, 1 << 20, d_osz, d_st, n, stream=s)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts = np.sort(np.array(ts[20:])) * 1e6
        # the reference's compression.c on ONE host thread (what one PostgreSQL backend gets) on the same blocks
        rb, ro, rs = ref.pack(chunks)
        ref.decompress(methods, rb, ro, rs, nthreads=1)
        cpu1 = ref.decompress(methods, rb, ro, rs, nthreads=1, reps=3)[2] / 3 * 1e6
        cpuN = ref.decompress(methods, rb, ro, rs, nthreads=os.cpu_count(), reps=3)[2] / 3 * 1e6
        ok = bool((d_st == 0).all().item()) and all(np.array_equal(d_dst[i].cpu().numpy(), blocks[i % len(z)]) for i in range(min(n, 8)))
        print(f"{tag:9s} batch {n:4d} (lz4/zstd alternating, S/M/D kinds): p50 {ts[len(ts)//2]:8.1f} us  p99 {ts[int(len(ts)*0.99)]:8.1f} us  "
              f"{n*(1<<20)/ts[len(ts)//2]/1e3:8.1f} GB/s at p50  exact={ok}  | reference CPU 1 thread {cpu1:9.1f} us, {os.cpu_count()} threads {cpuN:9.1f} us", flush=True)

if __name__ == "__main__":
    main()
