"""Quick device-resident timing of the individual stages at a cohort shape (development aid, not the bench)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import exomedepth_b200 as edb
from exomedepth_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=256)
ap.add_argument("--bins", type=int, default=200_000)
ap.add_argument("--states", type=int, default=5)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--direct", action="store_true", help="also time the in-register emission kernel")
ap.add_argument("--gen", type=int, default=16, help="distinct synthetic samples generated on the host (tiled)")
ap.add_argument("--seg", type=int, default=-1, help="segmented sweep: -1 auto, 0 off, 1 on")
ap.add_argument("--seg-min", type=int, default=0)
ap.add_argument("--seg-warm", type=int, default=0)
ap.add_argument("--warps", type=int, default=0)
ap.add_argument("--lattice", type=int, default=0)
ap.add_argument("--ref-scale", type=float, default=1.0, help="scale the synthetic counts (test and reference) by this factor")
ap.add_argument("--emission-only", action="store_true")
ap.add_argument("--first", type=int, default=0, help="index of the first synthetic sample")
a = ap.parse_args()

edb.init(0)
print(edb.device_info())
t0 = time.time()
d = synth.cohort(min(a.gen, a.samples), n_bins=a.bins, first_sample=a.first)
print("synth", time.time() - t0)
reps = (a.samples + d["observed"].shape[0] - 1) // d["observed"].shape[0]
if a.ref_scale != 1.0:
    d["observed"] = (d["observed"] * a.ref_scale).astype(np.int32)
    d["reference"] = (d["reference"] * a.ref_scale).astype(np.int32)
obs = np.tile(d["observed"], (reps, 1))[:a.samples]
phi = np.tile(d["phi"], reps)[:a.samples]
exp = np.tile(d["expected"], reps)[:a.samples]
t0 = time.time()
co = edb.Cohort(d["offsets"], d["start"], d["end"], n_states=a.states)
print("cohort create (host table build)", time.time() - t0, "table MB", co.table_bytes() / 1e6)
co.set_option("segments", a.seg).set_option("seg_min", a.seg_min).set_option("seg_warm", a.seg_warm).set_option("sweep_warps", a.warps)
dev = torch.device("cuda:0")
S, nb, ns = a.states, co.n_bins, a.samples
obs_t = torch.from_numpy(obs).to(dev)
ref_t = torch.from_numpy(d["reference"]).to(dev)
phi_t = torch.from_numpy(phi).to(dev)
exp_t = torch.from_numpy(exp).to(dev)
nbp = (nb + 15) // 16 * 16                                  # the Viterbi reads whole 128-byte lines of the rows
ll = torch.empty((ns, S, nbp), dtype=torch.float64, device=dev)
path = torch.empty((ns, nbp), dtype=torch.int8, device=dev)
calls = torch.zeros((ns, 512, 4), dtype=torch.int32, device=dev)
ncalls = torch.zeros(ns, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, name, bytes_per_cell):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    cells = ns * nb
    print(json.dumps(dict(stage=name, ms=round(ms, 4), ms_all=[round(t, 3) for t in ts],
                          Gcells_per_s=round(cells / ms / 1e6, 3), GBps=round(cells * bytes_per_cell / ms / 1e6, 1))))
    return ms


B = 4 + 8 * S
timeit(lambda: co.run_device(obs_t, ref_t, phi_t, exp_t, ll, path, calls, ncalls, what=1, mode=_lib.EMISSION_TABLE), "emission_table", B)
ll_tab = ll.clone()
print("ll checksum", float(ll.nan_to_num().sum()), "nan", int(ll.isnan().sum()))
if a.emission_only:
    sys.exit(0)
if a.direct:
    timeit(lambda: co.run_device(obs_t, ref_t, phi_t, exp_t, ll, path, calls, ncalls, what=1, mode=_lib.EMISSION_DIRECT), "emission_direct", B)
    diff = (ll - ll_tab).abs()
    print("table vs direct: max abs", float(diff.max()), "max rel", float((diff / ll.abs().clamp_min(1e-3)).max()))
    co.run_device(obs_t, ref_t, phi_t, exp_t, ll, path, calls, ncalls, what=1, mode=_lib.EMISSION_TABLE)
timeit(lambda: co.run_device(obs_t, ref_t, phi_t, exp_t, ll, path, calls, ncalls, what=2), "viterbi", 8 * S + 1)
_lib.profile(True)
for _ in range(3):
    co.run_device(obs_t, ref_t, phi_t, exp_t, ll, path, calls, ncalls, what=2)
print("per kernel ms:", {k: round(v[1] / v[0], 4) for k, v in _lib.profile_read().items()})
_lib.profile(False)
timeit(lambda: co.run_device(obs_t, ref_t, phi_t, exp_t, ll, path, calls, ncalls, what=3), "emission+viterbi", B + 1)
print("segments", co.segment_stats())
print("ncalls", ncalls[:8].tolist(), "status", _lib.load().edb200_status(0))
