"""One batch through the zstd pipeline in a fresh process (the pipeline reads its environment knobs once per process):
400 frames -- sparse and medium blocks written by the reference's libzstd, and sparse blocks written by the GPU encoder,
whose 64 KiB zstd blocks make every position the early pass of the raw / RLE stage guesses wrong.  Exit status 0: every
block bit-exact, none decoded by the fallback.  Run by tests/test_gpu_zstd_decode.py under several knob settings."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from oracle import ref                                     # noqa: E402  (the checker)
from pg_cryogen_b200 import COMP_ZSTD, CRYO_BLCKSZ, CryoGPU  # noqa: E402
from pg_cryogen_b200 import blockgen as bg                 # noqa: E402
from gpu_util import decode_device, encode_device          # noqa: E402


def main() -> int:
    gpu = CryoGPU(0)
    blocks = np.stack([bg.make_block("S", "hex", 51), bg.make_block("S", "lowcard", 52), bg.make_block("M", "hex", 53),
                       bg.make_block("S", "hex", 54)])
    lib, _, _ = ref.compress(COMP_ZSTD, 1, blocks[:3], nthreads=4)
    own, st = encode_device(gpu, COMP_ZSTD, 1, blocks[3:])
    assert (st == 0).all()
    pool = [(lib[0], 0), (lib[1], 1), (own[0], 3), (lib[2], 2)]
    pick = [pool[k % 4] for k in range(400)]
    out, osz, st = decode_device(gpu, COMP_ZSTD, [c for c, _ in pick])
    assert (st == 0).all() and (osz == CRYO_BLCKSZ).all(), "status / size"
    for k, (_, i) in enumerate(pick):
        assert np.array_equal(out[k], blocks[i]), ("block differs", k)
    frames, fallback = gpu.zstd_pipeline_stats()
    assert frames == 400 and fallback == 0, (frames, fallback)
    gpu.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
