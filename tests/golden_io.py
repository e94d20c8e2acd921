"""Reader for tests/golden/golden_v1.npz (written by tests/golden/make_golden.py from the compiled reference)."""
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def load_cases():
    z = np.load(PATH)
    index = json.loads(bytes(z["index"]).decode())
    cases = []
    for i, meta in enumerate(index):
        c = dict(meta)
        for k in ("cbegin", "points", "orders", "colors", "xf", "out", "stencil"):
            c[k] = z[f"c{i}_{k}"]
        cases.append(c)
    return cases


SCANLINE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scanline_v1.npz")


def load_scanline_cases():
    """tests/golden/scanline_v1.npz (written by tests/golden/make_scanline_golden.py from the compiled reference)."""
    z = np.load(SCANLINE_PATH)
    index = json.loads(bytes(z["index"]).decode())
    cases = []
    for i, meta in enumerate(index):
        c = dict(meta)
        for k in ("cbegin", "points", "orders", "colors", "xf", "out"):
            c[k] = z[f"c{i}_{k}"]
        if meta["kind"] == "sign":
            c["field"] = z[f"c{i}_field"]
        cases.append(c)
    return cases
