"""Pretty-print the headline numbers of bench.py JSON lines read from stdin."""
import json
import sys

for line in sys.stdin:
    try:
        d = json.loads(line)
    except Exception:
        print(line.rstrip())
        continue
    r = d.get("roofline", {})
    print(f"value {d.get('value', 0):.0f} {d.get('unit')}  ms/step {d.get('ms_per_step', 0):.2f}  "
          f"kernel {r.get('avg_launch_us', 0):.1f} us  {r.get('achieved', 0):.0f} GB/s  frac {r.get('frac', 0):.3f}  "
          f"e2e {d.get('e2e', {}).get('value', 0):.0f}  cpu {d.get('cpu_baseline', {}).get('value')}  clocks {d.get('clocks')}")
