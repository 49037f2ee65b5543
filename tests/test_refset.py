"""select.reference.set correlation sweep (R/optimize_reference_set.R:81-102).

CPU: the oracle restatement against the definition and against the independently written host mirror of the bin
filter; the sharded sweep over two gloo ranks (numpy stand-ins for the two compute stages).
GPU (-m gpu): the CUDA sweep through the C ABI against the oracle, tolerance 1e-10 on the correlations.
Parity unpinned: R is not available and the reference ships no fixture for this function (oracle/refset.py)."""
import multiprocessing as mp
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import refset as oref


def _cohort(ns=12, nb=6000, seed=1):
    from exomedepth_b200 import synth
    d = synth.cohort(ns, n_bins=nb)
    bl = (d["end"] - d["start"] + 1).astype(float)
    return d["observed"].astype(np.int32), bl


def test_bin_filter_mirror_matches_the_oracle():
    from exomedepth_b200 import refset
    counts, bl = _cohort()
    total = counts.sum(0)
    for kw in (dict(), dict(bin_length=bl), dict(bin_length=bl, n_bins_reduced=1000), dict(n_bins_reduced=777)):
        a = oref.select_bins(total, **kw)
        b = refset.select_bins(total, **kw)
        assert np.array_equal(a, b) and a.size > 100
        if kw.get("n_bins_reduced"):
            assert abs(a.size - kw["n_bins_reduced"]) <= 1 and np.all(np.diff(a) > 0)
    with pytest.raises(ValueError):
        refset.select_bins(total, bin_length=np.where(np.arange(total.size) == 5, 0.0, bl))


def test_oracle_correlations_are_the_pearson_matrix_of_length_scaled_counts():
    counts, bl = _cohort()
    sel, cor = oref.cohort_correlations(counts, bl)
    want = np.corrcoef(counts[:, sel] / bl[sel])             # the 1e6 / sum(x) factors cancel in a correlation
    np.testing.assert_allclose(cor, want, rtol=0, atol=1e-13)
    one = oref.correlations(counts[0, sel], counts[1:, sel].T, bl[sel])
    np.testing.assert_allclose(one, cor[0, 1:], rtol=0, atol=1e-13)
    assert oref.ranking(cor[0], 0)[0] == 1 + int(np.argmax(one))


def test_reference_doc_example_shape(exomecount):
    """R/optimize_reference_set.R:41-48: Exome1[1:200] against Exome2..4 — the front half runs and ranks 3 candidates."""
    ec = exomecount
    test = ec["Exome1"][:200]
    refs = np.stack([ec["Exome2"][:200], ec["Exome3"][:200], ec["Exome4"][:200]], 1)
    sel = oref.select_bins(refs.sum(1) + test)
    cor = oref.correlations(test[sel], refs[sel])
    assert 20 < sel.size < 200 and cor.shape == (3,) and np.all(cor > 0.5), (sel.size, cor)


# ---- two gloo ranks -------------------------------------------------------------------------------------------
class _NumpyBackend:
    @staticmethod
    def kpad(k):
        return (k + 15) // 16 * 16

    @staticmethod
    def standardize(counts, sel, bin_length):
        bl = np.ones(counts.shape[1]) if bin_length is None else np.asarray(bin_length, float)
        z = np.zeros((counts.shape[0], _NumpyBackend.kpad(sel.size)))
        for s in range(counts.shape[0]):
            y = oref.normalised(counts[s, sel], bl[sel])
            y = y - y.mean()
            z[s, :sel.size] = y / np.sqrt(np.sum(y * y))
        return z

    @staticmethod
    def gram(za, zb):
        return np.clip(za @ zb.T, -1, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, n_total, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    import torch.distributed as dist

    from exomedepth_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts, bl = _cohort(ns=n_total)
    lo, hi = shard.shard_range(n_total, rank, world)
    sel, cor = shard.refset_sweep(counts[lo:hi], n_total, bl, 1500, dist, backend=_NumpyBackend)
    q.put((rank, lo, hi, sel, cor))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 2])
def test_sharded_sweep_over_two_ranks_matches_single_process(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    counts, bl = _cohort(ns=n_total)
    sel, want = oref.cohort_correlations(counts, bl, 1500)
    for rank, lo, hi, s, cor in got:
        assert np.array_equal(s, sel) and cor.shape == (hi - lo, n_total)
        np.testing.assert_allclose(cor, want[lo:hi], rtol=0, atol=1e-12)


# ---- CUDA path -------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("reduced", [0, 5000])
def test_cuda_sweep_vs_oracle(reduced):
    from exomedepth_b200 import refset
    counts, bl = _cohort(ns=40, nb=20000)
    got = refset.cohort_reference_ranking(counts, bl, reduced)
    sel, want = oref.cohort_correlations(counts, bl, reduced)
    assert np.array_equal(got["selected"], sel)
    np.testing.assert_allclose(got["correlations"], want, rtol=1e-10, atol=1e-12)
    assert np.all(np.diag(got["correlations"]) == 1.0) or np.allclose(np.diag(got["correlations"]), 1.0, atol=1e-15)
    assert np.array_equal(got["correlations"], got["correlations"].T) or np.allclose(got["correlations"], got["correlations"].T, atol=1e-15)
    for t in range(0, 40, 7):
        ref_order = oref.ranking(want[t], t)
        gaps = np.abs(np.diff(want[t, ref_order]))
        if np.all(gaps > 1e-9):
            assert np.array_equal(got["order"][t], ref_order)
    # rows of a block against all — the block form the sharded sweep uses — are the same bits as in the full matrix
    blk = refset.correlations(counts, sel, bl, row0=13, n_rows=9)
    assert np.array_equal(blk, got["correlations"][13:22])


@pytest.mark.gpu
def test_cuda_select_reference_set_front_half(exomecount):
    from exomedepth_b200 import refset
    ec = exomecount
    test = ec["Exome1"][:4000]
    refs = np.stack([ec["Exome2"][:4000], ec["Exome3"][:4000], ec["Exome4"][:4000]], 1)
    got = refset.rank_candidates(test, refs, names=["Ex1", "Ex2", "Ex3"])
    sel = oref.select_bins(refs.sum(1) + test)
    cor = oref.correlations(test[sel], refs[sel])
    order = np.argsort(-cor, kind="stable")
    assert got["ref_samples"] == [["Ex1", "Ex2", "Ex3"][i] for i in order]
    np.testing.assert_allclose(got["correlations"], cor[order], rtol=1e-10)
    with pytest.raises(ValueError):
        refset.rank_candidates(test, refs[:-1])
    assert refset.rank_candidates(np.zeros(50, np.int32), refs[:50])["correlations"] is None      # :55-60


@pytest.mark.gpu
def test_cuda_select_reference_set_whole_function():
    """Correlation ranking, per-prefix beta-binomial fits, expected BF and the chosen prefix against the oracle
    restatement (scipy stand-ins for aod / VGAM: parity unpinned; tolerances are those of two optimisers)."""
    from exomedepth_b200 import refset, synth
    d = synth.cohort(9, n_bins=12000)
    test, refs = d["observed"][0], d["observed"][1:].T
    bl = (d["end"] - d["start"] + 1).astype(float)
    names = [f"S{i}" for i in range(1, 9)]
    got = refset.select_reference_set(test, refs, bl, names=names, chunk=3)
    want = oref.select_reference_set(test, refs, bl, names=names)
    st = got["summary_stats"]
    assert st["ref_samples"] == want["ref_samples"] and got["reference_choice"] == want["reference_choice"]
    assert 1 <= len(got["reference_choice"]) <= 8 and int(st["selected"].sum()) == 1
    np.testing.assert_allclose(st["correlations"], want["correlations"], rtol=1e-10)
    filled = ~np.isnan(want["phi"])
    assert np.array_equal(filled, ~np.isnan(st["phi"]))
    np.testing.assert_allclose(st["phi"][filled], want["phi"][filled], rtol=1e-5)
    np.testing.assert_allclose(st["mean_p"][filled], want["mean_p"][filled], rtol=1e-6)
    assert np.array_equal(st["median_depth"][filled], want["median_depth"][filled])
    np.testing.assert_allclose(st["RatioSd"][filled], want["RatioSd"][filled], rtol=1e-6)
    bf = ~np.isnan(want["expected_BF"])
    np.testing.assert_allclose(st["expected_BF"][bf], want["expected_BF"][bf], rtol=1e-4)


@pytest.mark.gpu
def test_cuda_device_stages_match_the_host_call():
    import torch
    from exomedepth_b200 import refset, shard
    counts, bl = _cohort(ns=21, nb=9000)
    sel, cor = shard.refset_sweep(counts, 21, bl, 0, None, device=torch.device("cuda:0"))
    want = refset.cohort_reference_ranking(counts, bl)
    assert np.array_equal(sel, want["selected"]) and np.array_equal(cor, want["correlations"])


# ---- two ranks on the GPU box: the sharded sweep with the exchange fused into the Gram kernel (CUDA IPC) ----------------
def _gpu_worker(rank, world, port_no, n_total, nccl, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    import torch
    import torch.distributed as dist

    import exomedepth_b200 as edb
    from exomedepth_b200 import shard, synth
    local = rank if nccl else 0                                 # one GPU: both ranks share it (IPC works within a device too)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    edb.init(local)
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    d = synth.cohort(8, n_bins=12000)
    rng = np.random.default_rng(3)
    counts = np.stack([rng.binomial(d["observed"][s % 8], 0.5 + 0.5 * rng.random()) for s in range(n_total)]).astype(np.int32)
    bl = (d["end"] - d["start"] + 1).astype(float)
    lo, hi = shard.shard_range(n_total, rank, world)
    out = {}
    # the all-gather form and the block-wise broadcasts need NCCL (one GPU per rank)
    forms = dict(fused=dict(fused=True))
    if nccl:
        forms.update(all_gather=dict(fused=False, blocks=False), blocks=dict(fused=False, blocks=True))
    for name, kw in forms.items():
        sel, cor = shard.refset_sweep(counts[lo:hi], n_total, bl, 0, dist, device=dev, **kw)
        out[name] = cor
    blocks = [None] * world
    dist.all_gather_object(blocks, (lo, out))
    if rank == 0:
        q.put((counts, sel, bl, sorted(blocks, key=lambda b: b[0])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_cuda_sharded_sweep_two_ranks_fused_exchange():
    """The multi-GPU form of the sweep inside `pytest -m gpu`: two processes, every rank standardises its block, the Gram
    kernel reads the other rank's rows through CUDA IPC (no all-gather pass); with two GPUs also the NCCL all-gather form
    and the block-wise broadcasts overlapped with the Gram kernel.
    Either way the assembled matrix equals the single-process matrix bit for bit (K-slices depend on K only)."""
    import torch
    import torch.multiprocessing as mp

    from exomedepth_b200 import refset
    nccl = torch.cuda.device_count() >= 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    n_total = 37
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port_no, n_total, nccl, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, sel, bl, blocks = q.get(timeout=600)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = refset.correlations(counts, sel, bl)
    for fused in blocks[0][1]:
        full = np.vstack([b[1][fused] for b in blocks])
        assert np.array_equal(full, want), fused
