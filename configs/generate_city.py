#!/usr/bin/env python
"""Seeded synthetic-city generator in the reference's buildings.xml schema.

The reference's config/generate_building.py:4-45 draws, per house and in this order, x, y, z, _R, _H with random.uniform
from Python's global (unseeded) generator.  Here the same draw order runs on random.Random(seed): generate_city(seed) is
exactly what the reference script produces after `random.seed(seed)` (pinned by tests/golden/city_golden.npz, recorded by
executing the reference script), so scaling runs over larger or denser cities are reproducible.

    python configs/generate_city.py --seed 7 --houses 26 --out configs/buildings_seed7.xml
"""
import argparse
import random

import numpy as np


def generate_city(seed, num_houses=26, x_range=(0, 500), y_range=(0, 500), z_range=(0, 0), r_range=(10, 50), h_range=(10, 50)):
    """-> float64 [num_houses, 5] = cx, cy, cz, _R, _H (the table engine.City / uavrl_env_config.buildings_host take)."""
    rng = random.Random(seed)
    out = np.zeros((num_houses, 5), np.float64)
    for i in range(num_houses):
        for j, rg in enumerate((x_range, y_range, z_range, r_range, h_range)):
            out[i, j] = rng.uniform(*rg)
    return out


def write_buildings_xml(table, path):
    """The reference's schema: <buildings><Threaten><Threaten_Type>building</Threaten_Type><position><x/><y/><z/></position>
    <_R/><_H/></Threaten>...; str() of a float round-trips exactly, as in the reference script."""
    with open(path, "w") as f:
        f.write("<?xml version='1.0' encoding='utf-8'?>\n<buildings>")
        for cx, cy, cz, R, H in np.asarray(table, np.float64):
            f.write("<Threaten><Threaten_Type>building</Threaten_Type><position><x>%s</x><y>%s</y><z>%s</z></position><_R>%s</_R><_H>%s</_H></Threaten>"
                    % (str(float(cx)), str(float(cy)), str(float(cz)), str(float(R)), str(float(H))))
        f.write("</buildings>")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, required=True)
    ap.add_argument("--houses", type=int, default=26)
    ap.add_argument("--size", type=float, default=500.0)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    t = generate_city(a.seed, a.houses, (0, a.size), (0, a.size))
    write_buildings_xml(t, a.out)
    print("wrote %d cylinders to %s" % (len(t), a.out))
