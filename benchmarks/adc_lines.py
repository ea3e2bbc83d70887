"""stdin: output of `python bench.py --workload adc` -> one short line per configuration."""
import json
import sys

for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        print(d["m"], round(d["ms"], 3), d.get("ms_each_call"), d.get("phases_ms"), "exact", d.get("exact_vs_reference"),
              "== scan", d.get("equal_to_lookup_scan"), "scan ms", round(d.get("lookup_scan_ms", 0), 2))
