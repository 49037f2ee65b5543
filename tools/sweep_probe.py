"""Device-resident timing of the Viterbi stage under different placements (development aid, not the bench).
  python tools/sweep_probe.py --opts sweep=2,vsplit=0,sweep_warps=2 --opts sweep=1 ...
Each --opts set is applied to a fresh cohort over the same emission matrix; prints per-kernel ms of the Viterbi pass."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import exomedepth_b200 as edb
from exomedepth_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=256)
ap.add_argument("--bins", type=int, default=200_000)
ap.add_argument("--states", type=int, default=5)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--gen", type=int, default=16)
ap.add_argument("--what", type=int, default=2, help="2 = Viterbi only over a resident matrix, 3 = emission + Viterbi")
ap.add_argument("--timeline", action="store_true")
ap.add_argument("--opts", action="append", default=[])
a = ap.parse_args()

edb.init(0)
d = synth.cohort(min(a.gen, a.samples), n_bins=a.bins)
reps = (a.samples + d["observed"].shape[0] - 1) // d["observed"].shape[0]
obs = np.tile(d["observed"], (reps, 1))[:a.samples]
phi = np.tile(d["phi"], reps)[:a.samples]
exp = np.tile(d["expected"], reps)[:a.samples]
dev = torch.device("cuda:0")
S, ns = a.states, a.samples
obs_t = torch.from_numpy(obs).to(dev)
ref_t = torch.from_numpy(d["reference"]).to(dev)
phi_t, exp_t = torch.from_numpy(phi).to(dev), torch.from_numpy(exp).to(dev)
nb = int(d["start"].size)
nbp = (nb + 15) // 16 * 16
ll = torch.empty((ns, S, nbp), dtype=torch.float64, device=dev)
path = torch.empty((ns, nbp), dtype=torch.int8, device=dev)
calls = torch.zeros((ns, 512, 4), dtype=torch.int32, device=dev)
ncalls = torch.zeros(ns, dtype=torch.int32, device=dev)
want = None
for spec in a.opts or [""]:
    co = edb.Cohort(d["offsets"], d["start"], d["end"], n_states=S)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        co.set_option(k, int(v))
    run = lambda what: co.run_device(obs_t, ref_t, phi_t, exp_t, ll, path, calls, ncalls, what=what)
    run(3)
    torch.cuda.synchronize()
    _lib.profile(2 if a.timeline else 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        run(a.what)
    e1.record()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile(0)
    got = (path.clone(), ncalls.clone())
    same = "" if want is None else f" identical={bool(torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]))}"
    want = want or got
    print(f"[{spec or 'default'}] {e0.elapsed_time(e1) / a.reps:.3f} ms/pass  " +
          "  ".join(f"{k}:{v[0] // a.reps}x avg {v[1] / v[0]:.3f} max {v[2]:.3f}" for k, v in prof.items()) + f"  calls {int(ncalls.sum())}{same}", flush=True)
    co.close()
