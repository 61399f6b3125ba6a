#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 block codec for pg_cryogen.

Workload (BASELINE.json configs[1], SURVEY.md 8(d) "Config 2"): the cryo blocks of a
synthetic 1M-row heap-tuple table (S blocks: 290 rows of 61 bytes per 1 MiB block,
3 449 blocks, 3.37 GiB of plaintext), compressed with zstd level 1 by the reference's
library, batch-decompressed on one B200 (the seq-scan read path, cache.c:178).

One "step" = one batched decompression of every block of the table.

  value      uncompressed GB/s with inputs resident in HBM (CUDA events on the launch stream)
  e2e        the same metric through the host-buffer C ABI call (cryogpu_decompress_host):
             pinned host buffers, H2D + kernels + D2H inside the timed region
  roofline   algorithmic bytes (csize_i + 1 MiB per block) / kernel time vs measured HBM peak
  cpu_baseline   the reference's own compression.c (oracle/_ref) on the box's host cores

N > 1 (torchrun, one rank per GPU): every rank decompresses its own table of the same
shape (weak scaling, no collective on the data path).

`--impl reference` times the reference's CPU implementation on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CRYO_BLCKSZ = 1 << 20
COMP_LZ4, COMP_ZSTD = 0, 1
NROWS = 1_000_000
KIND, PAYLOAD = "S", "hex"
METHOD, LEVEL = COMP_ZSTD, 1
METRIC = "decompress_GBps_zstd1_1Mrow_table"
# kernels of ours per step: method check, LZ4 decoder (exits: no LZ4 blocks), the eight pipeline stages
# (k_zp_sequences has two size classes), the fallback decoder (exits: nothing flagged)
LAUNCHES_PER_STEP = 11


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per step (all kernels of the zstd pipeline) from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("zstd_pipeline_dram_bytes_per_step")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_run(chunks, threads: int, target_seconds: float = 12.0):
    """Time the reference's compression.c (oracle/_ref) decompressing a bounded sample of the
    workload on the host cores.  Returns (GB/s, blocks in the sample, reps)."""
    from oracle import ref           # cpu_baseline leg: the one place bench.py touches oracle/
    sample = chunks[: min(len(chunks), max(64, threads * 16))]
    buf, offs, sizes = ref.pack(sample)
    out = np.empty((len(sample), CRYO_BLCKSZ), dtype=np.uint8)
    methods = np.full(len(sample), METHOD, dtype=np.int32)
    _, ok, t1 = ref.decompress(methods, buf, offs, sizes, nthreads=threads, reps=1, out=out)
    assert ok.all()
    reps = max(1, min(200, int(target_seconds / max(t1, 1e-4))))
    _, ok, t = ref.decompress(methods, buf, offs, sizes, nthreads=threads, reps=reps, out=out)
    gbs = len(sample) * reps * CRYO_BLCKSZ / t / 1e9
    return gbs, len(sample), reps


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    import benchdata
    threads = os.cpu_count() or 1
    nblk = min(512, 3449)
    chunks, _ = benchdata.build_table(NROWS, KIND, PAYLOAD, METHOD, LEVEL, count=nblk,
                                      threads=min(threads, 16))
    from oracle import ref
    buf, offs, sizes = ref.pack(chunks)
    out = np.empty((nblk, CRYO_BLCKSZ), dtype=np.uint8)
    methods = np.full(nblk, METHOD, dtype=np.int32)
    for _ in range(max(args.warmup, 1)):
        ref.decompress(methods, buf, offs, sizes, nthreads=threads, out=out)
    times = []
    for _ in range(args.steps):
        _, ok, t = ref.decompress(methods, buf, offs, sizes, nthreads=threads, out=out)
        assert ok.all()
        times.append(t)
    ms = 1e3 * sum(times) / len(times)
    gbs = nblk * CRYO_BLCKSZ / (ms * 1e-3) / 1e9
    sample = f"{nblk} of the 3449 zstd-1 S blocks per step, all {threads} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "1M-row S/hex table, zstd level 1, batched decompress (bounded sample)",
                   "blocks_per_step": nblk, "block_bytes": CRYO_BLCKSZ},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cryogpu", choices=["cryogpu", "reference"])
    ap.add_argument("--rows", type=int, default=NROWS, help="table rows (default: the 1M-row config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cryogpu" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import benchdata
    from pg_cryogen_b200 import CryoGPU, blockgen as bg, shard
    from pg_cryogen_b200.codec import pack_chunks

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libcryogpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload: this rank's table (weak scaling: same shape, different block seeds) ----
    nblk = bg.table_block_count(args.rows, KIND)
    host_threads = max(1, (os.cpu_count() or 8) // max(world, 1))
    chunks, plain = benchdata.build_table(args.rows, KIND, PAYLOAD, METHOD, LEVEL,
                                          threads=min(host_threads, 16),
                                          block_seed_offset=shard.rank_block_seed_offset(rank, nblk))
    buf, offs, sizes = pack_chunks(chunks)
    csize_total = int(sizes.astype(np.int64).sum())
    # host threads of the library's result placement (sparse return): this rank's share of cores
    os.environ.setdefault("CRYOGPU_HOST_THREADS", str(max(1, min(16, (os.cpu_count() or 8) // max(world, 1)))))
    gpu = CryoGPU(local_rank)
    d_src = torch.from_numpy(buf).to(dev)
    d_off = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_sz = torch.from_numpy(sizes.view(np.int32)).to(dev)
    d_me = torch.full((nblk,), METHOD, dtype=torch.int32, device=dev)
    d_dst = torch.empty((nblk, CRYO_BLCKSZ), dtype=torch.uint8, device=dev)
    d_osz = torch.zeros((nblk,), dtype=torch.int32, device=dev)
    d_st = torch.full((nblk,), -1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def step():
        gpu.decompress_device(d_me, d_src, d_off, d_sz, d_dst, CRYO_BLCKSZ, d_osz, d_st, nblk,
                              stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    # correctness gate: a wrong result is not a benchmark result
    assert bool((d_st == 0).all().item()), "decode status != 0"
    assert bool((d_osz == CRYO_BLCKSZ).all().item()), "decoded size != 1 MiB"
    assert np.array_equal(d_dst[: plain.shape[0]].cpu().numpy(), plain), "decoded bytes differ"

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    # the timed region is a few milliseconds, nvidia-smi samples every 100 ms: keep the same
    # steps running (untimed) under the sampler until it has seen the clocks under this load
    t_more = time.perf_counter()
    while len(sampler.rows) < 4 and time.perf_counter() - t_more < 3.0:
        for _ in range(20):
            step()
        torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    clocks["note"] = "sampled every 100 ms over the timed steps and the same steps repeated after them"
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = shard.max_over_ranks(total_ms, dev)          # slowest rank, on-device events
    ms_per_step = total_ms / args.steps
    value = shard.whole_job_rate(nblk * CRYO_BLCKSZ, world, ms_per_step * 1e-3) / 1e9

    # ---- roofline of the step: one batched decompression = the launches of the zstd pipeline
    #      (zstd_decode_p.cuh: parse, Huffman tables, literal streams, FSE tables, sequence walk
    #      x2 size classes, raw/RLE blocks, executor) plus three launches that exit at once.  The
    #      algorithmic bytes are the step's, so the duration is the step's too (CUDA events on the
    #      launching stream; the side streams are joined into it before the step ends). ----
    peak, peak_src = measured_peak()
    kern_ms = statistics.mean(step_ms)
    alg_bytes = nblk * CRYO_BLCKSZ + csize_total
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    frames, fallback = gpu.zstd_pipeline_stats()
    assert frames == nblk and fallback == 0, "the zstd pipeline handed frames to the fallback decoder"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": recorded_traffic(), "peak_source": peak_src,
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "kernel": "zstd pipeline (k_zp_parse, k_zp_huftab, k_zp_literals, k_zp_fsetab, "
                          "k_zp_sequences_small/large, k_zp_prefill, k_zp_execute); by device time "
                          "k_zp_execute and k_zp_prefill dominate (profiles/)",
                "algorithmic_bytes_per_launch": alg_bytes,
                "frames_decoded_by_fallback_kernel": fallback}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        lib = gpu.lib
        in_bytes = int(buf.size)
        h_in = lib.cryogpu_host_alloc(in_bytes)
        h_out = lib.cryogpu_host_alloc(nblk * CRYO_BLCKSZ)
        if not h_in or not h_out:
            raise SystemExit("pinned allocation failed")
        C.memmove(h_in, buf.ctypes.data, in_bytes)
        srcp = (C.c_void_p * nblk)(*[h_in + int(o) for o in offs])
        dstp = (C.c_void_p * nblk)(*[h_out + i * CRYO_BLCKSZ for i in range(nblk)])
        methods = np.full(nblk, METHOD, dtype=np.int32)
        osz = np.zeros(nblk, dtype=np.uint32)
        st = np.full(nblk, -1, dtype=np.int32)

        def host_step():
            rc = lib.cryogpu_decompress_host(gpu.handle, nblk, methods.ctypes.data, srcp,
                                             sizes.ctypes.data, dstp, CRYO_BLCKSZ,
                                             osz.ctypes.data, st.ctypes.data)
            if rc != 0:
                raise SystemExit("cryogpu_decompress_host: " + lib.cryogpu_last_error().decode())

        host_step()
        assert (st == 0).all()
        got = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_uint8)),
                                    shape=(plain.shape[0] * CRYO_BLCKSZ,))
        assert np.array_equal(got.reshape(plain.shape), plain), "e2e bytes differ"
        e2e_steps = max(1, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()
        torch.cuda.synchronize(dev)
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        t_e2e = shard.max_over_ranks(t_e2e, dev)
        bi, bo = gpu.last_transfer_bytes()              # counted by the library from what it copied
        e2e = {"value": shard.whole_job_rate(nblk * CRYO_BLCKSZ, world, t_e2e) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo,
               "steps": e2e_steps,
               "api": "cryogpu_decompress_host, pinned host buffers; only the non-zero 4 KiB pages of "
                      "each decoded block cross the bus, the library zero-fills the rest of the "
                      "caller's block with %s host threads" % os.environ["CRYOGPU_HOST_THREADS"]}
        lib.cryogpu_host_free(h_in)
        lib.cryogpu_host_free(h_out)

    # ---- CPU baseline beside it: the reference's compression.c on this box's cores ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        gbs_all, nsample, reps = cpu_reference_run(chunks, threads, 10.0)
        gbs_one, _, _ = cpu_reference_run(chunks, 1, 4.0)
        cpu = {"value": gbs_all, "unit": "GB/s", "cores": threads, "kind": "reference",
               "sample": f"{nsample} of the {nblk} blocks x {reps} passes through oracle/_ref "
                         f"(reference compression.c + libzstd 1.5.5), {threads} threads",
               "value_1_thread": gbs_one}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.rows}-row S/hex table ({nblk} cryo blocks of 1 MiB per GPU), "
                                   "zstd level 1 frames written by libzstd 1.5.5, batched decompress",
                       "blocks_per_gpu": nblk, "block_bytes": CRYO_BLCKSZ,
                       "compressed_bytes_per_gpu": csize_total,
                       "l2": "each step writes %.2f GB per GPU, far above the 126 MB L2"
                             % (nblk * CRYO_BLCKSZ / 1e9),
                       "parallelism": f"block-range shards x{world}, no collective"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": LAUNCHES_PER_STEP * args.steps, "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
