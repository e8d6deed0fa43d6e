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
