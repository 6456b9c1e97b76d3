"""Shared helpers for the parity tests (golden loading, key-sorted comparisons)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_config(g):
    return json.loads(str(g["config_json"]))


def sort_by_key(keys, *vals):
    """Lexicographic (i, j, k) order; returns (keys_sorted, *vals_sorted)."""
    keys = np.asarray(keys, dtype=np.int64).reshape(-1, 3)
    order = np.lexsort(keys.T[::-1])
    return (keys[order],) + tuple(np.asarray(v)[order] for v in vals)


def assert_same_map(keys_a, L_a, keys_b, L_b, atol=1e-5, what=""):
    """Bit-exact key sets, log-odds within atol (north_star: 1e-5 absolute)."""
    ka, la = sort_by_key(keys_a, L_a)
    kb, lb = sort_by_key(keys_b, L_b)
    assert ka.shape == kb.shape, f"{what}: {len(ka)} vs {len(kb)} voxels"
    assert np.array_equal(ka, kb), f"{what}: voxel key sets differ"
    err = np.abs(la - lb).max() if len(la) else 0.0
    assert err <= atol, f"{what}: max |dL| = {err}"
    return err


def world_keys_of_points(points, resolution):
    return np.floor(np.asarray(points, dtype=np.float64) / resolution).astype(np.int64)
