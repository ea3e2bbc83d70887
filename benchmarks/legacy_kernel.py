#!/usr/bin/env python
"""Shim: `python benchmarks/legacy_kernel.py ARGS` == `python bench.py --workload legacy ARGS` (the reference's own CUDA kernel on the same GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.execv(sys.executable, [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "legacy"] + sys.argv[1:])
