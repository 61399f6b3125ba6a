"""Load tests/golden/reference_streams.npz -> list of (name, method, level, stream, sha256)."""
import os

import numpy as np

from pg_cryogen_b200 import blockgen as bg

HERE = os.path.dirname(os.path.abspath(__file__))

# how to rebuild the plaintext of each fixture (must match tests/golden/make_golden.py)
PLAINTEXT = {
    "regr_rows_1_290": lambda: bg.regression_block(1, 290),
    "regr_rows_291_500": lambda: bg.regression_block(291, 500),
    "regr_rows_501_790": lambda: bg.regression_block(501, 790),
    "regr_rows_791_1000": lambda: bg.regression_block(791, 1000),
    "S_hex_b7": lambda: bg.make_block("S", "hex", 7),
    "S_lowcard_b7": lambda: bg.make_block("S", "lowcard", 7),
    "M_lowcard_b7": lambda: bg.make_block("M", "lowcard", 7),
    "D_lowcard_b7": lambda: bg.make_block("D", "lowcard", 7),
}


def load():
    z = np.load(os.path.join(HERE, "golden", "reference_streams.npz"))
    names = sorted(k[: -len("__stream")] for k in z.files if k.endswith("__stream"))
    out = []
    for n in names:
        method, level = (int(v) for v in z[n + "__method"])
        out.append((n, method, level, z[n + "__stream"], bytes(z[n + "__sha256"])))
    return out


def plaintext(name: str) -> np.ndarray:
    key = name.rsplit("_", 1)[0]
    return PLAINTEXT[key]()
