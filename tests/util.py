"""Seeded synthetic problems shared by the tests, the golden-vector script and bench.py."""
import numpy as np


def sift_like(rng, n, d):
    """SIFT-shaped data: non-negative small integers stored as float32."""
    return np.clip(np.floor(np.abs(rng.standard_normal((n, d))) * 40.0), 0, 218).astype(np.float32)


def make_problem(seed, n, d, m, h=256, kind="sift"):
    """-> X (n, d) f32, C (m, h, d) f32, B (n, m) int16 1-based."""
    rng = np.random.default_rng(seed)
    X = sift_like(rng, n, d) if kind == "sift" else rng.standard_normal((n, d)).astype(np.float32)
    pool = sift_like(rng, m * h, d) if kind == "sift" else rng.standard_normal((m * h, d)).astype(np.float32)
    C = (pool.reshape(m, h, d) / np.float32(m)).astype(np.float32)
    if kind == "sift":
        C = np.round(C * 4) / 4  # quarter-integers: exact in fp32, makes ties plausible
    C = C.astype(np.float32)
    B = rng.integers(1, h + 1, size=(n, m)).astype(np.int16)
    return X, C, B


def make_scan_problem(seed, n, nq, d, m, h=256):
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, h, size=(n, m)).astype(np.uint8)
    queries = sift_like(rng, nq, d)
    codebooks = (sift_like(rng, m * h, d) / np.float32(m)).astype(np.float32)
    recon_norm = rng.random(n).astype(np.float32) * 1000.0
    return codes, queries, codebooks, recon_norm
