import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
scans, _ = bench.make_scans(106, 42, "cuda:0")
print(json.dumps(bench.dropin_leg(scans, 100, 5, show_stderr=True)))
