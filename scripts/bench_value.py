"""Reads one bench.py JSON line on stdin, prints a short summary (label = argv[1])."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1] if len(sys.argv) > 1 else "", f"value={d['value']:.4e}", f"ms/step={d['ms_per_step']:.2f}",
      f"e2e={d['e2e']['value']:.3e}" if d.get('e2e', {}).get('value') else "", f"frac={d['roofline']['frac']:.4f}")
