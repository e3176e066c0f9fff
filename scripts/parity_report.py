"""Full-size parity of the BASELINE configs C3 and C2 against the oracle, every unit compared (run under gpurun; the
summary printed here is what profiles/r02_parity_full.json holds)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ila_b200
import parity_full

out = {"c3": parity_full.c3_full_report(ila_b200), "c2": parity_full.c2_full_report(ila_b200)}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_parity_full.json"), "w"), indent=1)
sys.exit(0 if out["c3"]["pass"] and out["c2"]["pass"] else 1)
