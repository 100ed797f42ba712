"""Runs the FP64 / FP32 peak micro-benchmarks (roofline denominators) alone; capture with
ncu --set full -k regex:k_peak_fp64 -c 1 to show the pipe utilisation of the measurement itself."""
import sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from lqr_obstacles_b200 import Lqro
ctx = Lqro(n_agents=2, max_neighbors=1, horizon=20)
print('fp64, fp32 TFLOP/s', ctx.measure_peaks())
