"""Device-resident timing of the reference-set correlation sweep (development aid): standardise + Gram kernels."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import exomedepth_b200 as edb
from exomedepth_b200 import _lib, refset, synth

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=256)
ap.add_argument("--rows", type=int, default=0, help="rows of the matrix this GPU forms (0 = all)")
ap.add_argument("--bins", type=int, default=200_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cpu", type=int, default=0, help="also time numpy on this many samples")
a = ap.parse_args()
edb.init(0)
d = synth.cohort(16, n_bins=a.bins)
rng = np.random.default_rng(0)
base = d["observed"]
counts = np.empty((a.samples, base.shape[1]), np.int32)
for s in range(a.samples):                                  # distinct samples: thin a base sample binomially
    counts[s] = rng.binomial(base[s % 16], rng.uniform(0.6, 1.0))
bl = (d["end"] - d["start"] + 1).astype(float)
sel = refset.select_bins(counts.sum(0, dtype=np.int64), bl)
dev = torch.device("cuda:0")
c_t, sel_t, bl_t = torch.from_numpy(counts).to(dev), torch.from_numpy(sel).to(dev), torch.from_numpy(bl).to(dev)
kp = refset.kpad(sel.size)
z = torch.empty((a.samples, kp), dtype=torch.float64, device=dev)
m = a.rows or a.samples
out = torch.empty((m, a.samples), dtype=torch.float64, device=dev)
print(f"samples {a.samples} rows {m} bins {counts.shape[1]} selected {sel.size}")


def run():
    refset.standardize_device(c_t, sel_t, bl_t, z)
    refset.gram_device(z[:m], z, sel.size, out)


run()
torch.cuda.synchronize()
_lib.profile(True)
for _ in range(a.reps):
    run()
prof = _lib.profile_read()
_lib.profile(False)
for k, v in prof.items():
    print(f"  {k:20s} {v[1] / v[0]:9.4f} ms")
g = prof["refset_gram"][1] / prof["refset_gram"][0]
print(f"  gram: {2.0 * m * a.samples * kp / g / 1e9:.2f} TFLOP/s FP64; standardise: "
      f"{(2 * 4 * a.samples * sel.size + 8 * a.samples * kp) / (prof['refset_standardize'][1] / prof['refset_standardize'][0]) / 1e6:.1f} GB/s")
if a.cpu:
    from oracle import refset as oref
    t0 = time.perf_counter()
    _, want = oref.cohort_correlations(counts[:a.cpu], bl)
    print(f"  numpy oracle on {a.cpu} samples: {time.perf_counter() - t0:.3f} s")
# the beta-binomial fit of every sample against the shared reference aggregate
ref_t = torch.from_numpy(d["reference"]).to(dev)
mu = torch.empty(a.samples, dtype=torch.float64, device=dev)
phi = torch.empty_like(mu)
ll = torch.empty_like(mu)
info = torch.empty(a.samples, dtype=torch.int32, device=dev)
L = _lib.load()
st = torch.cuda.current_stream().cuda_stream
_lib.profile(True)
for _ in range(a.reps):
    _lib.check(L.edb200_betabin_fit_device(c_t.data_ptr(), c_t.stride(0), ref_t.data_ptr(), 0, a.samples, c_t.shape[1],
                                           mu.data_ptr(), phi.data_ptr(), ll.data_ptr(), info.data_ptr(), st), "fit")
prof = _lib.profile_read()
_lib.profile(False)
print(f"  betabin_fit          {prof['betabin_fit'][1] / prof['betabin_fit'][0]:9.4f} ms for {a.samples} samples; "
      f"iterations min/median/max {int(info.min())}/{int(info.median())}/{int(info.max())}")
