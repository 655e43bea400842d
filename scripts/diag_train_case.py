"""Diagnostic (GPU box): per-parameter error of a training fixture case on cuda."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import deepflows_b200
import parity
from conftest import rel_err
from DeepFlows import backend_api
name = sys.argv[1] if len(sys.argv) > 1 else "resnet_registered"
backend_api.set_precision("fp32"); backend_api.set_dgrad_mode("reference")
res, g = parity.run_training_case(name, "cuda")
print("losses", res["losses"], g["losses"])
ill = set(g.get("ill_conditioned", np.array([], dtype="U1")).tolist())
for k, v in res["params"].items():
    e = rel_err(v, g["p1." + k])
    d = np.abs(v - g["p1." + k])
    upd = np.abs(g["p1." + k] - g["p0." + k]).max()
    print("%-26s rel %.3e  abs %.3e  max|p| %.3e  max|update| %.3e %s" % (k, e, d.max(), np.abs(g["p1." + k]).max(), upd, "ILL" if k in ill else ""))
for k, v in res["stats"].items():
    if k in g: print("%-26s rel %.3e" % (k, rel_err(v, g[k])))
