"""Two device-resident steps of the small-panel shape (BASELINE.json configs[3]: 512 samples x 5,000 bins x 7 states) —
a short target for ncu captures of its kernels (development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import exomedepth_b200 as edb
from exomedepth_b200 import synth

S, ns = int(os.environ.get("PANEL_STATES", 7)), 512
edb.init(0)
dev = torch.device("cuda", 0)
d = synth.cohort(64, per_chrom=(5, 1000))
nb, nbp = 5000, 5008
co = edb.Cohort(d["offsets"], d["start"], d["end"], n_states=S)
obs = torch.from_numpy(np.tile(d["observed"], (ns // 64, 1))).to(dev)
ref = torch.from_numpy(d["reference"]).to(dev)
phi, exp = torch.from_numpy(np.tile(d["phi"], ns // 64)).to(dev), torch.from_numpy(np.tile(d["expected"], ns // 64)).to(dev)
ll = torch.empty((ns, S, nbp), dtype=torch.float64, device=dev)
path = torch.empty((ns, nbp), dtype=torch.int8, device=dev)
calls = torch.zeros((ns, 256, 4), dtype=torch.int32, device=dev)
ncalls = torch.zeros(ns, dtype=torch.int32, device=dev)
for _ in range(2):
    co.run_device(obs, ref, phi, exp, ll, path, calls, ncalls, what=3)
torch.cuda.synchronize()
print("calls", int(ncalls.sum()))
