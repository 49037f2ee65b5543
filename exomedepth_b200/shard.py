"""Multi-GPU plumbing: samples shard across ranks, the shared exon-bin metadata is broadcast once.

The hot path has no data-path collective: every sample is independent for the emission, every
(sample, chromosome) chain is independent for the Viterbi.  One process per GPU; rank 0 builds the shared bin
geometry, the reference aggregate and — on the host, with the host libm, so that every rank holds the same bits —
the log-transition table, and broadcasts them (NCCL on GPUs, gloo in the CPU tests).  Results stay sharded; only
the small call tables are gathered.  torch.distributed is plumbing here, nothing more.

The reference-set correlation sweep (select.reference.set) is the one place with a real
exchange: the bin filter needs the cohort-wide total per bin (one all-reduce of a per-rank partial sum) and every
rank needs every sample's standardised row (one all-gather) to form its block of the correlation matrix.
"""
import numpy as np


def shard_range(n_samples, rank, world):
    """Contiguous block of ceil(n/world) samples for `rank` (the last ranks may get fewer, or none)."""
    per = -(-n_samples // world)
    lo = min(rank * per, n_samples)
    return lo, min(lo + per, n_samples)


def broadcast_arrays(arrays, names, dist=None, src=0, device=None):
    """Broadcast a dict of numpy arrays from rank `src`.  `arrays` may be None on the other ranks; `names` is the
    ordered key list every rank knows.  One size/dtype header, then one payload per array."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return {k: np.ascontiguousarray(arrays[k]) for k in names}
    import torch
    rank = dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    codes = {"int32": 0, "int64": 1, "float64": 2, "int8": 3, "uint8": 4}
    inv = {v: k for k, v in codes.items()}
    hdr = torch.zeros((len(names), 4), dtype=torch.int64)
    if rank == src:
        for i, k in enumerate(names):
            a = np.ascontiguousarray(arrays[k])
            assert a.ndim <= 2, k
            shape = list(a.shape) + [1] * (2 - a.ndim)
            hdr[i] = torch.tensor([codes[a.dtype.name], a.ndim, shape[0], shape[1]])
    hdr = hdr.to(dev)
    dist.broadcast(hdr, src)
    hdr = hdr.cpu().numpy()
    out = {}
    for i, k in enumerate(names):
        dt = np.dtype(inv[int(hdr[i, 0])])
        ndim = int(hdr[i, 1])
        shape = tuple(int(x) for x in hdr[i, 2:2 + ndim]) if ndim else ()
        n = int(np.prod(shape)) if ndim else 1
        if rank == src:
            buf = torch.from_numpy(np.ascontiguousarray(arrays[k]).reshape(-1).view(np.uint8).copy()).to(dev)
        else:
            buf = torch.empty(n * dt.itemsize, dtype=torch.uint8, device=dev)
        dist.broadcast(buf, src)
        out[k] = buf.cpu().numpy().view(dt).reshape(shape).copy()
    return out


def gather_calls(calls, ncalls, first_sample, dist=None, dst=0, device=None):
    """Gather the per-sample call tables of every rank on `dst`.

    calls: int32[n_local, cap, 4], ncalls: int32[n_local]; first_sample: global index of this rank's first sample.
    Returns on `dst` a list of (global sample index, int32[n, 4]) sorted by sample, elsewhere None."""
    local = [(first_sample + s, np.asarray(calls[s, :min(int(ncalls[s]), calls.shape[1])], np.int32)) for s in range(len(ncalls))]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    flat = np.concatenate([np.concatenate([[g, c.shape[0]], c.reshape(-1)]) for g, c in local]).astype(np.int64) if local else np.zeros(0, np.int64)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([flat.size], dtype=torch.int64, device=dev))
    most = max(int(s.item()) for s in sizes)
    pad = torch.zeros(max(most, 1), dtype=torch.int64, device=dev)
    pad[:flat.size] = torch.from_numpy(flat).to(dev)
    bufs = [torch.zeros(max(most, 1), dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        v = bufs[r].cpu().numpy()[:int(sizes[r].item())]
        i = 0
        while i < v.size:
            g, n = int(v[i]), int(v[i + 1])
            out.append((g, v[i + 2:i + 2 + 4 * n].reshape(n, 4).astype(np.int32)))
            i += 2 + 4 * n
    return sorted(out, key=lambda t: t[0])


def make_cohort(shared, dist=None, device=None, **cohort_kwargs):
    """Build this rank's Cohort from the broadcast shared metadata.  Rank 0 builds the host-libm log-transition
    table; the other ranks receive its bytes (one broadcast), so every rank runs the Viterbi on identical terms."""
    import torch

    from .cohort import Cohort
    multi = dist is not None and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if multi else 0
    co = Cohort(shared["offsets"], shared["start"], shared["end"], build_table=(rank == 0), **cohort_kwargs)
    if multi:
        tbl = torch.empty(co.table_bytes() // 8, dtype=torch.float64, device=device)
        if rank == 0:
            co.table_to(tbl)
        torch.cuda.synchronize()
        dist.broadcast(tbl, 0)
        if rank != 0:
            co.table_from(tbl)
        torch.cuda.synchronize()
    return co


def exchange_and_gram(z_local, n_local, per, n_total, n_selected, dist, dev, bufs=None, outs=None):
    """The exchange of the sharded reference-set sweep overlapped with its contraction: one asynchronous NCCL broadcast
    per rank's block of standardised rows (z_local: [per, k_pad], this rank's block, zero-padded), and this rank's first
    n_local rows against block j as soon as it has arrived.  Returns the list of [n_local, rows of block j] correlation
    blocks (bufs / outs: reusable receive buffers and outputs)."""
    import torch
    from . import refset
    world, rank = dist.get_world_size(), dist.get_rank()
    if bufs is None:
        bufs = [z_local if j == rank else torch.empty_like(z_local) for j in range(world)]
    rows = [max(0, min(per, n_total - j * per)) for j in range(world)]
    if outs is None:
        outs = [torch.empty((n_local, rows[j]), dtype=torch.float64, device=dev) for j in range(world)]
    works = [dist.broadcast(bufs[j], src=j, async_op=True) for j in range(world)]
    for j in range(world):
        works[j].wait()                                     # the compute stream waits for block j, not the host
        if n_local and rows[j]:
            refset.gram_device(z_local[:n_local], bufs[j][:rows[j]], n_selected, outs[j])
    return outs


def refset_sweep(counts_local, n_total, bin_length=None, n_bins_reduced=0, dist=None, device=None, backend=None, fused=None, blocks=None):
    """Sharded leave-one-out correlation sweep.  counts_local: int32[n_local, n_bins] — this rank's samples
    (contiguous blocks as shard_range deals them); n_total: samples over all ranks.
    Returns (selected bins, float64[n_local, n_total] correlations of this rank's samples against every sample).

    Collectives: all-reduce (sum) of the per-bin totals, all-gather of the standardised rows.  `backend` supplies the
    two compute stages — default: the CUDA kernels (exomedepth_b200.refset); the CPU tests pass numpy stand-ins.
    fused=True (CUDA, one process per GPU of one NVLink box): no all-gather — every rank maps the other ranks' blocks
    with CUDA IPC and the Gram kernel reads its B tiles from their owners' memory (same bits as the all-gather form).
    fused=None picks it when a rank forms at most 256 rows AND there are at most 4 ranks (two row tiles: every remote tile
    then crosses NVLink at most twice; measured: 2 GPUs 2.6 vs 2.9 ms at 256 rows per rank and 35 vs 33 ms at 1,000; 4 GPUs
    5.2 vs 5.4 ms at 250 rows; 8 GPUs 11.7 vs 11.3 ms at 250 rows — with seven peers the remote reads of the Gram kernel
    cost more than NCCL's all-gather, profiles/r2f_bench_refset_n8*.json).  blocks=True (default beyond two ranks when not
    fused): the exchange as one broadcast per block, overlapped with the Gram kernel block by block (exchange_and_gram)."""
    from . import refset
    multi = dist is not None and dist.is_initialized() and dist.get_world_size() > 1
    world = dist.get_world_size() if multi else 1
    rank = dist.get_rank() if multi else 0
    counts_local = np.ascontiguousarray(np.asarray(counts_local, np.int32))
    n_local, n_bins = counts_local.shape
    import torch
    dev = device if device is not None else torch.device("cpu")
    total = torch.from_numpy(counts_local.sum(0, dtype=np.int64)).to(dev)
    if multi:
        dist.all_reduce(total)
    sel = refset.select_bins(total.cpu().numpy(), bin_length, n_bins_reduced)
    per = -(-n_total // world)                              # every rank contributes a block of `per` rows (zero padded)
    if fused is None:
        fused = per <= 256 and world <= 4 and not blocks
    if blocks is None:
        blocks = not fused and world > 2                    # (two ranks: one remote block, nothing to overlap it with but the local one)
    if fused and multi and backend is None:
        c_t = torch.from_numpy(counts_local).to(dev)
        sel_t = torch.from_numpy(sel).to(dev)
        bl_t = None if bin_length is None else torch.from_numpy(np.ascontiguousarray(np.asarray(bin_length, np.float64))).to(dev)
        z_ptr, handle = refset.block_alloc(per, sel.size)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        refset.peers_open(handles, rank)
        if n_local:
            refset.standardize_device(c_t, sel_t, bl_t, z_ptr)
        torch.cuda.synchronize()
        dist.barrier()                                      # every block is complete before anyone reads it
        out = torch.empty((n_local, n_total), dtype=torch.float64, device=dev)
        if n_local:
            refset.gram_peers_device(n_local, per, n_total, sel.size, out)
        torch.cuda.synchronize()
        dist.barrier()                                      # nobody unmaps a block that is still being read
        refset.peers_close()
        return sel, out.cpu().numpy()
    if backend is None:
        kp = refset.kpad(sel.size)
        c_t = torch.from_numpy(counts_local).to(dev)
        sel_t = torch.from_numpy(sel).to(dev)
        bl_t = None if bin_length is None else torch.from_numpy(np.ascontiguousarray(np.asarray(bin_length, np.float64))).to(dev)
        z_local = torch.zeros((per, kp), dtype=torch.float64, device=dev)
        if n_local:
            refset.standardize_device(c_t, sel_t, bl_t, z_local[:n_local])
    else:
        z = backend.standardize(counts_local, sel, bin_length)
        z_local = torch.zeros((per, z.shape[1] if n_local else backend.kpad(sel.size)), dtype=torch.float64, device=dev)
        if n_local:
            z_local[:n_local] = torch.from_numpy(z).to(dev)
    if multi and backend is None and blocks:
        # block by block: every rank's block is broadcast in turn (NCCL, asynchronous) and this rank's rows are contracted
        # with block j while block j + 1 is still crossing NVLink — the all-gather no longer stands in front of the Gram
        # kernel.  Same bits as the other forms: an entry's K-slices do not depend on how the columns are grouped.
        out_blocks = exchange_and_gram(z_local, n_local, per, n_total, sel.size, dist, dev)
        out = torch.cat(out_blocks, dim=1) if n_local else torch.empty((0, n_total), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        return sel, out.cpu().numpy()
    if multi:
        z_all = torch.empty((world * per, z_local.shape[1]), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(z_all, z_local)
        z_all = z_all[:n_total]                             # only the last ranks' blocks are padded
    else:
        z_all = z_local[:n_total]
    if backend is None:
        out = torch.empty((n_local, n_total), dtype=torch.float64, device=dev)
        if n_local:
            refset.gram_device(z_local[:n_local], z_all.contiguous(), sel.size, out)
            torch.cuda.synchronize()
        return sel, out.cpu().numpy()
    return sel, backend.gram(z_local[:n_local].cpu().numpy(), z_all.cpu().numpy())
