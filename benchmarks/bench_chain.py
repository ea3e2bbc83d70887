#!/usr/bin/env python
"""Shim: `python benchmarks/bench_chain.py ARGS` == `python bench.py --workload chain ARGS` (chain / Viterbi encoder benchmark)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.execv(sys.executable, [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "chain"] + sys.argv[1:])
