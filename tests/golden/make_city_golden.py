"""Record what the reference's config/generate_building.py writes after random.seed(s) (the script itself is unseeded):
tests/golden/city_golden.npz pins configs/generate_city.py.  Run here (needs /root/reference): python tests/golden/make_city_golden.py"""
import importlib.util
import os
import random
import sys
import tempfile
import xml.etree.ElementTree as ET

import numpy as np

REF = "/root/reference/config/generate_building.py"
spec = importlib.util.spec_from_file_location("ref_generate_building", REF)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)                                  # the __main__ block does not run on import
out = {}
for seed, n in ((7, 26), (2024, 60)):
    random.seed(seed)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "b.xml")
        mod.generate_houses_xml(n, (0, 500), (0, 500), (0, 0), (10, 50), (10, 50), path)   # the script's own __main__ arguments
        rows = []
        for t in ET.parse(path).getroot().findall("Threaten"):
            p = t.find("position")
            rows.append([float(p.find(k).text) for k in ("x", "y", "z")] + [float(t.find("_R").text), float(t.find("_H").text)])
    out["seed%d_n%d" % (seed, n)] = np.array(rows)
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "city_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
