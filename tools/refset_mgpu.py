"""Sharded reference-set sweep over NCCL (run under torchrun, one rank per GPU): every rank standardises its samples,
the rows are all-gathered, every rank forms its block of the correlation matrix; rank 0 checks the assembled matrix
against the single-GPU call and prints device times (max over ranks)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

import exomedepth_b200 as edb
from exomedepth_b200 import refset, shard, synth

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=512)
ap.add_argument("--bins", type=int, default=200_000)
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--fused", type=int, default=-1, help="1: no all-gather, the Gram kernel reads the peers' blocks over NVLink (CUDA IPC); "
                                                     "0: NCCL all-gather, then Gram; -1: the library's choice (fused up to 256 rows per rank)")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
edb.init(local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"                  # the version banner goes to stdout
    dist.init_process_group("nccl", device_id=dev)
d = synth.cohort(16, n_bins=a.bins)
rng = np.random.default_rng(0)                              # every rank generates the same cohort, then keeps its block
counts = np.empty((a.samples, d["observed"].shape[1]), np.int32)
for s in range(a.samples):
    counts[s] = rng.binomial(d["observed"][s % 16], rng.uniform(0.6, 1.0))
bl = (d["end"] - d["start"] + 1).astype(float)
lo, hi = shard.shard_range(a.samples, rank, world)
grp = dist if world > 1 else None
shard.refset_sweep(counts[lo:hi], a.samples, bl, 0, grp, device=dev, fused=None if a.fused < 0 else bool(a.fused))          # warm-up
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0.record()
sel, cor = shard.refset_sweep(counts[lo:hi], a.samples, bl, 0, grp, device=dev, fused=None if a.fused < 0 else bool(a.fused))
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
blocks = [None] * world
if world > 1:
    dist.all_gather_object(blocks, (lo, hi, cor))
else:
    blocks = [(lo, hi, cor)]
if rank == 0:
    full = np.vstack([b[2] for b in sorted(blocks, key=lambda b: b[0])])
    ok = None
    if a.check:
        want = refset.correlations(counts, sel, bl)
        ok = bool(np.array_equal(full, want))
        print("max |difference| to the single-GPU matrix:", float(np.max(np.abs(full - want))))
    print(json.dumps(dict(workload=f"reference-set sweep, {a.samples} samples x {counts.shape[1]} bins, {sel.size} selected bins",
                          n_gpus=world, fused={-1: "auto", 0: False, 1: True}[a.fused], ms_sweep_incl_upload_and_collectives=float(t[0]), identical_to_single_gpu=ok)))
# ---- device time of the exchange + contraction stage alone (inputs resident, rows already standardised) ----------
if world > 1:
    sel_t = torch.from_numpy(sel).to(dev)
    bl_t = torch.from_numpy(bl).to(dev)
    c_t = torch.from_numpy(counts[lo:hi]).to(dev)
    per = -(-a.samples // world)
    kp = refset.kpad(sel.size)
    n_local = hi - lo
    out = torch.empty((n_local, a.samples), dtype=torch.float64, device=dev)

    def stage_time(fn, reps=5):
        best = 1e30
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t[0]))
        return best

    z_local = torch.zeros((per, kp), dtype=torch.float64, device=dev)
    refset.standardize_device(c_t, sel_t, bl_t, z_local[:n_local])
    z_all = torch.empty((world * per, kp), dtype=torch.float64, device=dev)

    def gather_then_gram():
        dist.all_gather_into_tensor(z_all, z_local)
        refset.gram_device(z_local[:n_local], z_all[:a.samples], sel.size, out)

    t_nccl = stage_time(gather_then_gram)
    ref_out = out.clone()
    z_ptr, handle = refset.block_alloc(per, sel.size)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    refset.peers_open(handles, rank)
    refset.standardize_device(c_t, sel_t, bl_t, z_ptr)
    torch.cuda.synchronize()
    dist.barrier()
    t_fused = stage_time(lambda: refset.gram_peers_device(n_local, per, a.samples, sel.size, out))
    same = bool(torch.equal(out, ref_out))
    dist.barrier()
    refset.peers_close()
    if rank == 0:
        print(json.dumps(dict(stage="exchange + Gram, device time, max over ranks", n_gpus=world, samples=a.samples,
                              ms_nccl_all_gather_then_gram=t_nccl, ms_fused_peer_memory_gram=t_fused, identical=same,
                              all_gather_bytes_per_rank=int((world - 1) * per * kp * 8))))
    dist.destroy_process_group()
