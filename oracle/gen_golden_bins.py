"""Golden walk over the frames of a bin, produced by the REFERENCE's own decompress_array.  TEST INFRASTRUCTURE.

Run from the repo root (only where /root/reference exists):  python oracle/gen_golden_bins.py
Imports /root/reference/clair/utils.py with `blosc` replaced by a stub whose unpack_array is clair_b200.bins.unpack_array
(python-blosc is absent here; the frames are written by clair_b200.bins.pack_array), then records what the reference's
decompress_array (clair/utils.py:223-262) returns step by step for several frame layouts and batch sizes:
tests/golden/bins_walk.json.  What this pins is the walk (which rows, which next indices), not the frame codec.
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"
SCENARIOS = [  # (rows per frame, rows to retrieve per call)
    ([500, 500, 137], 1000), ([500, 500, 500, 500], 1000), ([500, 500, 137], 300), ([500, 500, 137], 500),
    ([500, 500, 137], 64), ([500], 1000), ([500, 1], 500), ([7, 7, 7, 3], 5), ([500, 500, 500, 20], 10000),
]


def frames_of(sizes):
    from clair_b200 import bins
    out, start = [], 0
    for n in sizes:
        out.append(bins.pack_array(np.arange(start, start + n, dtype=np.int32).reshape(n, 1)))
        start += n
    return out


def walk(decompress_array, frames, batch, read_index_list=None):
    steps, index, first = [], 0, 0
    for _ in range(200):
        rows, first_next, index_next = decompress_array(frames, index, first, batch, len(frames), read_index_list)
        steps.append([None if rows is None else [int(len(rows)), int(rows[0, 0]), int(rows[-1, 0])], int(first_next), int(index_next)])
        if index_next == -1:
            break
        index, first = index_next, first_next
    return steps


def main():
    if not os.path.isdir(REFERENCE):
        print("no /root/reference here: bins_walk.json not regenerated")
        return
    from clair_b200 import bins
    sys.path.insert(0, REFERENCE)
    stub = types.ModuleType("blosc")
    stub.unpack_array = bins.unpack_array
    stub.pack_array = lambda a, **kw: bins.pack_array(a)
    stub.set_nthreads = lambda n: None
    stub.NOSHUFFLE = 0
    sys.modules["blosc"] = stub
    tree = types.ModuleType("intervaltree")
    tree.IntervalTree = object
    sys.modules.setdefault("intervaltree", tree)
    import clair.utils as ref_utils
    golden = []
    for sizes, batch in SCENARIOS:
        frames = frames_of(sizes)
        golden.append({"sizes": sizes, "batch": batch, "steps": walk(ref_utils.decompress_array, frames, batch)})
        order = list(range(len(sizes)))[::-1]
        golden.append({"sizes": sizes, "batch": batch, "read_index_list": order,
                       "steps": walk(ref_utils.decompress_array, frames, batch, order)})
    path = os.path.join(ROOT, "tests", "golden", "bins_walk.json")
    with open(path, "w") as f:
        json.dump(golden, f)
    print(path, "written:", len(golden), "walks,", sum(len(g["steps"]) for g in golden), "steps")


if __name__ == "__main__":
    main()
