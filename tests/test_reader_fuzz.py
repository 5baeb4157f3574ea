"""Robustness of the product's .vdb reader: truncated and bit-flipped files must end in a VdbError (or parse), never in
a crash or an unbounded allocation.  The reference panics on malformed input (unwrap / panic!); the
C ABI promises status codes instead."""
import numpy as np
import pytest

import vdb_writer as V
import woxel_b200 as W


def small_file(comp):
    rng = np.random.default_rng(5)
    pts = np.unique(np.concatenate([rng.integers(0, 16, size=(150, 3)), rng.integers(-16, 0, size=(60, 3))]), axis=0)
    return V.VdbWriter(compression=comp, half_float=True, leaf_metadata=6).build(pts), len(pts)


@pytest.mark.parametrize("comp", [V.ACTIVE_MASK, V.ZIP | V.ACTIVE_MASK, V.BLOSC | V.ACTIVE_MASK])
def test_mutated_files_never_crash(tmp_path, comp):
    good, n = small_file(comp)
    rng = np.random.default_rng(comp)
    p = tmp_path / "m.vdb"
    outcomes = {"ok": 0, "error": 0}
    # the node masks (2 N5 x 8 KB + ...) dominate the file: aim most mutations at the header and the tail (leaf data)
    spots = np.concatenate([rng.integers(0, 400, 60), rng.integers(len(good) - 3000, len(good), 120), rng.integers(0, len(good), 40)])
    for k, at in enumerate(spots.tolist()):
        b = bytearray(good)
        if k % 3 == 0:
            b = b[:at]  # truncation
        else:
            b[at] ^= 1 << (k % 8)
            if k % 5 == 0 and at + 4 < len(b):
                b[at:at + 4] = b"\xff\xff\xff\x7f"  # a huge length field
        p.write_bytes(bytes(b))
        try:
            v = W.VdbReader(str(p)).read_vdb345_grid("ls_test")
            v.count_leaf_values()
            outcomes["ok"] += 1
        except (W.vdb.VdbError, MemoryError):
            outcomes["error"] += 1
    assert outcomes["error"] > 20 and outcomes["ok"] + outcomes["error"] == len(spots)
