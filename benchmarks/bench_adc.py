#!/usr/bin/env python
"""Shim: `python benchmarks/bench_adc.py ARGS` == `python bench.py --workload adc ARGS` (ADC linear-scan benchmark)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.execv(sys.executable, [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "adc"] + sys.argv[1:])
