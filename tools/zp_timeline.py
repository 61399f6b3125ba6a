"""dev tool (GPU box): build libcryogpu with -DZP_TIMELINE into a scratch .so, run the headline batch,
print when each pipeline kernel started / ended (GPU global timer, microseconds from the first start)."""
import ctypes as C, os, subprocess, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.environ.get("TL_LIB") or os.path.join(ROOT, "tools", "_prof", "libcryogpu_tl.so")      # prebuilt in the build container, or built here
os.makedirs(os.path.dirname(so), exist_ok=True)
if not os.path.exists(so):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DZP_TIMELINE",
                           "-Xcompiler", "-fPIC", "-diag-suppress", "550", "-shared", "-o", so,
                           os.path.join(ROOT, "pg_cryogen_b200", "csrc", "cryogpu.cu")])
import benchdata
from pg_cryogen_b200 import CryoGPU, blockgen as bg, codec
from pg_cryogen_b200.codec import pack_chunks
codec.lib_path = lambda: so
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kind, payload = (sys.argv[2], sys.argv[3]) if len(sys.argv) > 3 else ("S", "hex")
nblk = bg.table_block_count(rows, kind)
chunks, plain = benchdata.build_table(rows, kind, payload, 1, 1, threads=16)
buf, offs, sizes = pack_chunks(chunks)
extra = [CryoGPU(0) for _ in range(int(os.environ.get("TL_EXTRA_CONTEXTS", "0")))]    # contexts made before the one measured
gpu = CryoGPU(0)
L = gpu.lib
assert hasattr(L, "cryogpu_debug_timeline"), "timeline build not loaded"
dev = torch.device("cuda:0")
d_src = torch.from_numpy(buf).to(dev); d_off = torch.from_numpy(offs.view(np.int64)).to(dev)
d_sz = torch.from_numpy(sizes.view(np.int32)).to(dev); d_me = torch.full((nblk,), 1, dtype=torch.int32, device=dev)
d_dst = torch.empty((nblk, 1 << 20), dtype=torch.uint8, device=dev)
d_osz = torch.zeros((nblk,), dtype=torch.int32, device=dev); d_st = torch.full((nblk,), -1, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
if os.environ.get("TL_LZ4_FIRST"):
    # the context's first batch is an LZ4 one (tools/gpu_probe.py order): the state in which the zstd step is slow
    from oracle import ref
    lb = bg.make_blocks("S", "hex", 0, 32)
    lc, _, _ = ref.compress(0, 1, lb, nthreads=8)
    lbuf, loffs, lsz = pack_chunks([lc[i % 32] for i in range(nblk)])
    l_src = torch.from_numpy(lbuf).to(dev); l_off = torch.from_numpy(loffs.view(np.int64)).to(dev)
    l_sz = torch.from_numpy(lsz.view(np.int32)).to(dev); l_me = torch.zeros((nblk,), dtype=torch.int32, device=dev)
    for _ in range(3):
        gpu.decompress_device(l_me, l_src, l_off, l_sz, d_dst, 1 << 20, d_osz, d_st, nblk, stream=s)
    torch.cuda.synchronize()
    del l_src
names = ["parse", "prefill", "huftab", "literals", "fsetab", "seq_small", "seq_large", "execute",
         "lz4_route", "lz4_warp", "lz4_cta", "execute_cta", "zstd_warp", "prefill_early"]
for it in range(4):
    L.cryogpu_debug_timeline(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gpu.decompress_device(d_me, d_src, d_off, d_sz, d_dst, 1 << 20, d_osz, d_st, nblk, stream=s)
    e1.record(); torch.cuda.synchronize()
    out = (C.c_ulonglong * 32)()
    L.cryogpu_debug_timeline(out, 0)
    t0 = min(out[2 * k] for k in range(len(names)) if out[2 * k + 1])
    print(f"iteration {it}: {e0.elapsed_time(e1) * 1e3:.0f} us by events")
    for k, nm in enumerate(names):
        if out[2 * k + 1]:
            print(f"   {nm:10s} start {(out[2*k]-t0)/1e3:8.1f}  end {(out[2*k+1]-t0)/1e3:8.1f}  ({(out[2*k+1]-out[2*k])/1e3:7.1f} us)")
assert bool((d_st == 0).all().item())
