"""Batched Riccati solve alone: 4,096 heterogeneous problems and a batch of one (bench.py's riccati arm).  usage: python tools/quick_riccati.py"""
import json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import bench
from lqr_obstacles_b200 import Lqro
ctx = Lqro(n_agents=64, max_neighbors=8, horizon=20)
from lqr_obstacles_b200.api import sim_default_model
ctx.sim_set_model(sim_default_model())
print(json.dumps(bench.riccati_arm(ctx)))
