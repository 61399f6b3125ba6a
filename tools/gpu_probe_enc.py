"""dev probe: time batched compression on the GPU and report the ratio vs the reference."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
sys.path.insert(0, ".")
from pg_cryogen_b200 import CryoGPU, blockgen as bg
from pg_cryogen_b200.codec import compress_bound
from oracle import ref

def probe(method, level, kind, payload, n):
    g = CryoGPU(0)
    uniq = min(n, 32)
    blocks = bg.make_blocks(kind, payload, 0, uniq)
    _, rsz, _ = ref.compress(method, level, blocks, nthreads=8)
    dev = torch.device("cuda:0")
    d_src = torch.from_numpy(blocks).to(dev).repeat((n + uniq - 1) // uniq, 1)[:n].contiguous()
    bound = compress_bound(method); stride = (bound + 15) & ~15
    d_dst = torch.zeros((n, stride), dtype=torch.uint8, device=dev)
    d_sz = torch.zeros((n,), dtype=torch.int32, device=dev); d_st = torch.full((n,), -1, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    def run():
        g.compress_device(method, level, d_src, 1 << 20, d_dst, stride, stride, d_sz, d_st, n, stream=s)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = int(os.environ.get("PROBE_REPS", "3"))
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ok = bool((d_st == 0).all().item())
    gsz = d_sz[:uniq].cpu().numpy()
    print(f"enc method={method} level={level} {kind}/{payload} n={n}: {ms:.3f} ms {n*(1<<20)/ms/1e6:.1f} GB/s in, "
          f"status_ok={ok} ratio gpu/ref size = {gsz.sum()/rsz.sum():.3f} (worst block {np.max(gsz/rsz):.3f})", flush=True)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    for sp in sys.argv[2:] or ["0:1:S:hex", "0:1:M:hex", "0:1:D:hex", "0:1:D:lowcard",
                              "1:1:S:hex", "1:1:M:hex", "1:1:D:hex", "1:1:D:lowcard", "1:-5:D:lowcard", "1:3:M:lowcard"]:
        m, l, k, pl = sp.split(":")
        try:
            probe(int(m), int(l), k, pl, n)
        except Exception as ex:
            print("probe failed", sp, ex, flush=True)
