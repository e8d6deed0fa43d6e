"""Host-side logic of the reference-facing plug-ins that needs no GPU: the XML loader reproduces the
xmltodict mapping the reference relies on, the shipped-format configs parse into the reference's
parameters, and the epsilon schedule matches simulator.epsilon_annealing (golden, simulator.py:141-145)."""
import math
import os

import numpy as np

import uavrl_b200  # noqa: F401
from uavrl_b200.plugins import xmlconfig
from uavrl_b200.plugins.PathPlan_City_B200 import buildings_from_dict, uav_params_from_dict
from conftest import ROOT

CFG = os.path.join(ROOT, "configs")


def test_xml2dict_mapping_and_city(env_golden):
    d = xmlconfig.XML2Dict(os.path.join(CFG, "PathPlan_City_B200.xml"))
    env = d["simulator"]["env"]
    assert env["Env_Type"] == "PathPlan_City_B200" and env["len"] == "500"          # leaves stay strings
    assert env["Agent"]["Trainer"]["Trainer_path"].endswith("Trainer_DQN_B200.xml")  # nested dicts
    b = xmlconfig.XML2Dict(os.path.join(CFG, "buildings.xml"))["buildings"]
    assert isinstance(b["Threaten"], list) and len(b["Threaten"]) == 26             # repeated tags -> list
    table = buildings_from_dict(b)
    assert np.array_equal(table, env_golden["buildings"])                           # doubles round-trip exactly
    assert xmlconfig.None2Value(None, 3) == 3 and xmlconfig.None2Value("x", 3) == "x"


def test_uav_parameters_follow_reference_casts(env_golden):
    uav = xmlconfig.XML2Dict(os.path.join(CFG, "UAV_B200.xml"))["Agent"]
    p = uav_params_from_dict(uav)
    g = env_golden["uav_params"]                     # Max_V, Min_V, Steering_angle (rad), Max_Step from the reference
    assert p.max_v == g[0] and p.min_v == g[1] and p.max_step == int(g[3])
    assert p.steering == g[2] == 30.0 / 180 * math.pi


def test_epsilon_annealing_kat(dqn_golden):
    for mx, mn, ep, want in dqn_golden["eps_schedule"]:
        assert xmlconfig.epsilon_annealing(ep, mn, mx) == want


def test_seeded_city_generator_matches_reference_script(tmp_path):
    """configs/generate_city.py == the reference's config/generate_building.py run after random.seed(seed) (golden recorded by
    executing the reference script), and its XML parses back to the same table through the plug-ins' loader."""
    import importlib.util
    import os
    import numpy as np
    from conftest import GOLDEN, ROOT
    from uavrl_b200.plugins import xmlconfig
    spec = importlib.util.spec_from_file_location("generate_city", os.path.join(ROOT, "configs", "generate_city.py"))
    gc = importlib.util.module_from_spec(spec); spec.loader.exec_module(gc)
    g = np.load(os.path.join(GOLDEN, "city_golden.npz"))
    for key in g.files:
        seed, n = int(key.split("_")[0][4:]), int(key.split("_n")[1])
        t = gc.generate_city(seed, n)
        assert np.array_equal(t, g[key]), key
        path = os.path.join(tmp_path, key + ".xml")
        gc.write_buildings_xml(t, path)
        th = xmlconfig.XML2Dict(path)["buildings"]["Threaten"]
        back = np.array([[float(x["position"]["x"]), float(x["position"]["y"]), float(x["position"]["z"]), float(x["_R"]), float(x["_H"])] for x in th])
        assert np.array_equal(back, t)
    assert (t[:, 3] >= 10).all() and (t[:, 3] <= 50).all() and (t[:, :2] >= 0).all() and (t[:, :2] <= 500).all()
