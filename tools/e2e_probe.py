"""End-to-end timing of the host-pointer cohort call (development aid, not the bench): pinned host buffers in,
CNV calls + per-call columns out; variants with the Viterbi path and the likelihood matrix copied back."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import exomedepth_b200 as edb
from exomedepth_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=256)
ap.add_argument("--bins", type=int, default=200_000)
ap.add_argument("--states", type=int, default=5)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--ll", action="store_true")
ap.add_argument("--i32", action="store_true", help="int32 counts instead of the 16-bit ingestion layout")
ap.add_argument("--p12", action="store_true", help="12-bit ingestion layout")
ap.add_argument("--timeline", action="store_true")
ap.add_argument("--opts", default="", help="cohort options, e.g. parts=4,sweep=1; several sets separated by ';' are timed in turn")
ap.add_argument("--calls-only", action="store_true")
a = ap.parse_args()
edb.init(0)
d = synth.cohort(16, n_bins=a.bins)
reps = (a.samples + 15) // 16
co = edb.Cohort(d["offsets"], d["start"], d["end"], n_states=a.states)
ns, nb, S, cap = a.samples, co.n_bins, a.states, 1024
hb = _lib.PinnedPool()
obs = hb.empty((ns, nb), np.int32)
obs[:] = np.tile(d["observed"], (reps, 1))[:ns]
ovf = None
if a.p12:
    obs, oi, ov = edb.pack_counts12(obs, out=hb.empty((ns, ((nb + 1) // 2 * 3 + 3) // 4 * 4), np.uint8))
    ovf = (oi, ov)
elif not a.i32:
    obs, oi, ov = edb.pack_counts(obs, out=hb.empty((ns, nb), np.uint16))
    ovf = (oi, ov)
phi = np.tile(d["phi"], reps)[:ns]
ex = np.tile(d["expected"], reps)[:ns]
out = dict(calls=hb.empty((ns, cap, 4), np.int32), ncalls=hb.empty((ns,), np.int32), call_stats=hb.empty((ns, cap, 3), np.float64),
           cor=hb.empty((ns,), np.float64), path=hb.empty((ns, nb), np.int8))
if a.ll:
    out["ll"] = hb.empty((ns, S, nb), np.float64)
variants = [("calls+stats", dict(want_ll=False, want_path=False, want_stats=True)),
            ("calls+stats+path", dict(want_ll=False, want_path=True, want_stats=True))]
if a.calls_only:
    variants = variants[:1]
if a.ll:
    variants.append(("calls+stats+path+ll", dict(want_ll=True, want_path=True, want_stats=True)))
for opts in a.opts.split(";"):
    for name in ("chunks", "reserve"):
        co.set_option(name, 0)
    for kv in filter(None, opts.split(",")):
        co.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    for name, kw in variants:
        kw["overflow"] = ovf
        co.run_host(obs, d["reference"], phi, ex, call_cap=cap, out=out, **kw)
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            r = co.run_host(obs, d["reference"], phi, ex, call_cap=cap, out=out, **kw)
            ts.append(1e3 * (time.perf_counter() - t0))
        if a.timeline:
            _lib.profile(2)
            co.run_host(obs, d["reference"], phi, ex, call_cap=cap, out=out, **kw)
            _lib.profile_read()
            _lib.profile(0)
        print(f"{name:22s} opts={opts or 'default'} {'p12' if a.p12 else 'i32' if a.i32 else 'u16'}: best {min(ts):.3f} ms  median {sorted(ts)[len(ts) // 2]:.3f}  "
              f"{ns * nb / min(ts) / 1e6:.2f} G bin*samples/s  calls {int(r['ncalls'].sum())}", flush=True)
