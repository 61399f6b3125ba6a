"""dev probe: latency of the call the reference makes (cache.c:178: ONE block per call) through
cryogpu_decompress_device, per method and block kind, beside the reference's compression.c on one host thread."""
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
    reps = int(os.environ.get("ONE_REPS", "120"))
    for method in (0, 1):
        for kind, pl in (("S", "hex"), ("S", "lowcard"), ("M", "hex"), ("M", "lowcard"), ("D", "hex"), ("D", "lowcard")):
            blk = bg.make_block(kind, pl, 5)[None]
            c = ref.compress(method, 1, blk)[0]
            buf, offs, sz = pack_chunks(c)
            d_src = torch.from_numpy(buf).to(dev); d_off = torch.from_numpy(offs.view(np.int64)).to(dev)
            d_sz = torch.from_numpy(sz.view(np.int32)).to(dev); d_me = torch.tensor([method], dtype=torch.int32, device=dev)
            d_dst = torch.empty((1, 1 << 20), dtype=torch.uint8, device=dev)
            d_osz = torch.zeros((1,), dtype=torch.int32, device=dev); d_st = torch.full((1,), -1, dtype=torch.int32, device=dev)
            s = torch.cuda.current_stream().cuda_stream
            ts = []
            for it in range(reps + 10):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                g.decompress_device(d_me, d_src, d_off, d_sz, d_dst, 1 << 20, d_osz, d_st, 1, stream=s)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            ts = np.sort(np.array(ts[10:])) * 1e6
            # the call the drop-in makes: host pointers, one block (cryogpu_decompress_host behind cryo_decompress)
            hs = []
            out = np.zeros((1, 1 << 20), dtype=np.uint8)
            for it in range(reps // 2 + 5):
                t0 = time.perf_counter()
                g.decompress_host([method], c, out=out)
                hs.append(time.perf_counter() - t0)
            hs = np.sort(np.array(hs[5:])) * 1e6
            okh = np.array_equal(out[0], blk[0])
            rb, ro, rs = ref.pack(c)
            ref.decompress([method], rb, ro, rs, nthreads=1)
            cpu1 = ref.decompress([method], rb, ro, rs, nthreads=1, reps=5)[2] / 5 * 1e6
            ok = int(d_st[0].item()) == 0 and np.array_equal(d_dst[0].cpu().numpy(), blk[0])
            print(f"one block {'lz4 ' if method == 0 else 'zstd'} {kind}/{pl:8s} csize {len(c[0]):7d}: GPU p50 {ts[len(ts)//2]:8.1f} us p99 {ts[int(len(ts)*0.99)]:8.1f} us"
                  f" | host-pointer call p50 {hs[len(hs)//2]:8.1f} us | reference CPU 1 thread {cpu1:8.1f} us | exact={ok and okh}", flush=True)

if __name__ == "__main__":
    main()
