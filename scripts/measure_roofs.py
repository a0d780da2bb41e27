#!/usr/bin/env python
"""Roofs of the box (run under gpurun): random-record gather bandwidth over an L2-resident and an HBM-resident table,
and the pinned host -> device ceiling with 1..N GPUs copying at once.  Prints one JSON object (kept under profiles/)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nimble_aligner_b200 as nb

out = {"gather_gbs": {}, "h2d": {}}
for name, tb in (("l2_24MB", 24 << 20), ("l2_47MB", 47 << 20), ("l2_96MB", 96 << 20), ("hbm_2GB", 2 << 30), ("hbm_10GB", 10 << 30)):
    for rec in (32, 64):
        out["gather_gbs"]["%s_rec%d" % (name, rec)] = round(nb.measure_gather(tb, rec), 1)
nd = nb.lib().nb_device_count()
n = 1
while n <= nd:
    per, agg = nb.measure_h2d(list(range(n)), 512 << 20, 8)
    out["h2d"]["n%d" % n] = {"aggregate_gbs": round(agg, 1), "per_device_gbs": [round(x, 1) for x in per]}
    n *= 2
print(json.dumps(out))
