"""Reads one bench.py JSON line on stdin and prints the headline + per-stage medians (us)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), 'frames/s', round(d['ms_per_step'] * 1e3, 1), 'us/step',
      {k: round(v['median'] * 1e3, 1) for k, v in d['stages_ms'].items()})
