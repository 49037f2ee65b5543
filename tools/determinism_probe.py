"""Repeat the host-pointer call and report whether every repetition returns the same bytes (development aid)."""
import argparse
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import exomedepth_b200 as edb
from exomedepth_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=256)
ap.add_argument("--reps", type=int, default=12)
ap.add_argument("--i32", action="store_true")
ap.add_argument("--path", type=int, default=0)
ap.add_argument("--pinned", type=int, default=1)
ap.add_argument("--opts", action="append", default=[])
a = ap.parse_args()
edb.init(0)
d = synth.cohort(16, n_bins=200_000)
reps = (a.samples + 15) // 16
obs = np.tile(d["observed"], (reps, 1))[:a.samples]
phi, ex = np.tile(d["phi"], reps)[:a.samples], np.tile(d["expected"], reps)[:a.samples]
ovf = None
if not a.i32:
    obs, oi, ov = edb.pack_counts(obs)
    ovf = (oi, ov)
hb = _lib.PinnedPool()
if a.pinned:
    o2 = hb.empty(obs.shape, obs.dtype)
    o2[:] = obs
    obs = o2
ns, nb = obs.shape
out = dict(calls=hb.empty((ns, 1024, 4), np.int32), ncalls=hb.empty((ns,), np.int32), call_stats=hb.empty((ns, 1024, 3), np.float64),
           cor=hb.empty((ns,), np.float64), path=hb.empty((ns, nb), np.int8)) if a.pinned else None
for spec in a.opts or [""]:
    co = edb.Cohort(d["offsets"], d["start"], d["end"], n_states=5)
    for kv in filter(None, spec.split(",")):
        co.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    seen = {}
    for r in range(a.reps):
        res = co.run_host(obs, d["reference"], phi, ex, call_cap=1024, want_ll=False, want_path=bool(a.path), want_stats=True, overflow=ovf, out=out)
        h = hashlib.sha256()
        n = res["ncalls"]
        for k in (("path",) if a.path else ()) + ("ncalls",):
            h.update(np.ascontiguousarray(res[k]).tobytes())
        for s_i in range(ns):
            h.update(np.ascontiguousarray(res["calls"][s_i, :n[s_i]]).tobytes())
        key = (h.hexdigest()[:12], int(res["ncalls"].sum()))
        if len(seen) and key[:2] not in seen:
            first = next(iter(seen.values()))[0]
            print("   differs at rep", r, "samples with different ncalls:", np.flatnonzero(n != keep_n)[:10], "ncalls", n[n != keep_n][:10], "vs", keep_n[n != keep_n][:10])
        if not seen:
            keep_n = n.copy()
        seen.setdefault(key[:2], []).append(r)
    print(f"[{spec or 'default'}] distinct results: {len(seen)}  {seen}", flush=True)
    co.close()
