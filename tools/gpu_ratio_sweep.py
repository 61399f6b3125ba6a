"""dev probe: worst-case gpu_csize / reference_csize per lz4_acceleration (1 MiB blocks, all kinds x payloads)
and per zstd level -- the data behind the stated ratio tolerance."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from pg_cryogen_b200 import CryoGPU, blockgen as bg
from gpu_util import encode_device
from oracle import ref

def main():
    g = CryoGPU(0)
    blocks, tags = [], []
    for kind in "SMD":
        for pl in bg.PAYLOADS:
            for seed in (11, 12):
                blocks.append(bg.make_block(kind, pl, seed)); tags.append(f"{kind}/{pl}")
    blocks.append(bg.regression_block(1, 290)); tags.append("regression")
    blocks = np.stack(blocks)
    for method, levels in ((0, [0, 1, 2, 3, 5, 8, 10, 15, 20, 25, 30, 40, 50]), (1, [-5, -4, -3, -2, -1, 0, 1, 2, 3])):
        for lv in levels:
            comp, st = encode_device(g, method, lv, blocks)
            assert (st == 0).all()
            _, rs, _ = ref.compress(method, lv, blocks, nthreads=8)
            r = np.array([len(c) for c in comp]) / rs
            w = int(np.argmax(r))
            print(f"method={method} level={lv:3d}: mean {r.mean():.3f} worst {r.max():.3f} ({tags[w]} gpu {len(comp[w])} ref {int(rs[w])})  "
                  f"abs worst excess {int((np.array([len(c) for c in comp]) - rs).max())} B", flush=True)

if __name__ == "__main__":
    main()
