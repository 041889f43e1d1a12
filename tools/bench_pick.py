"""Debug aid: print a few fields of a bench.py JSON line read from stdin."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1] if len(sys.argv) > 1 else "", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 4), round(d["streaming_ms_per_scan"], 4),
      [(c.get("driver", "threads")[:7], c["sequences"], round(c["scans_per_s"])) for c in (d.get("concurrent_sequences_one_gpu") or [])])
