"""N > 1 host logic on CPU: two gloo ranks shard the samples, receive the broadcast shared metadata and gather the
call tables; the union must equal the single-process result.  The per-sample arithmetic here is the oracle port
(this is a test of the sharding / collective plumbing of exomedepth_b200/shard.py, not of the CUDA path)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sample_calls(port, framing, shared, obs, phi, e, S):
    ll = port.emission(phi, e, obs + shared["reference"], obs, port.state_odds(S))
    T = port.callcnvs_transitions(S, 1e-4)
    off = shared["offsets"]
    rows = []
    for c in range(len(off) - 1):
        b0, b1 = int(off[c]), int(off[c + 1])
        loc, pos = framing.frame_chromosome(ll[b0:b1], shared["start"][b0:b1].astype(float), shared["end"][b0:b1].astype(float), 50000.0)
        for (sp, ep, typ, nex) in port.c_hmm(T, loc, pos, 50000.0)[1]:
            rows.append([sp - 1 + b0, ep - 1 + b0, typ, nex])
    return np.array(rows, np.int32).reshape(-1, 4)


def _worker(rank, world, port_no, n_samples, S, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    import torch.distributed as dist

    from exomedepth_b200 import shard, synth
    from oracle import framing, port
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = ["offsets", "start", "end", "reference"]
    arrays = None
    if rank == 0:
        off, start, end = synth.geometry(3000)
        arrays = dict(offsets=off, start=start, end=end, reference=synth.shared(start.size)[1])
    shared = shard.broadcast_arrays(arrays, names, dist)
    lo, hi = shard.shard_range(n_samples, rank, world)
    cap = 64
    calls = np.zeros((hi - lo, cap, 4), np.int32)
    ncalls = np.zeros(hi - lo, np.int32)
    for i, s in enumerate(range(lo, hi)):
        obs, phi, e = synth.sample(s, shared["reference"], n_segments=12)
        rows = _sample_calls(port, framing, shared, obs, phi, e, S)
        ncalls[i] = rows.shape[0]
        calls[i, :rows.shape[0]] = rows[:cap]
    got = shard.gather_calls(calls, ncalls, lo, dist)
    if rank == 0:
        q.put((shared, got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [5, 2, 1])
def test_two_rank_sharding_matches_single_process(n_samples):
    import torch.multiprocessing as mp

    from exomedepth_b200 import shard, synth
    from oracle import framing, port
    port.lib()
    S, world = 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pn = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, pn, n_samples, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    shared, got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # shared metadata arrived intact
    off, start, end = synth.geometry(3000)
    assert np.array_equal(shared["offsets"], off) and np.array_equal(shared["start"], start) and np.array_equal(shared["end"], end)
    assert shared["start"].dtype == np.int32 and shared["offsets"].dtype == np.int64
    # every sample exactly once, in order, with the single-process calls
    assert [g for g, _ in got] == list(range(n_samples))
    ref = synth.shared(start.size)[1]
    sh = dict(offsets=off, start=start, end=end, reference=ref)
    for g, c in got:
        obs, phi, e = synth.sample(g, ref, n_segments=12)
        assert np.array_equal(c, _sample_calls(port, framing, sh, obs, phi, e, S))


def test_shard_ranges_cover_every_sample_once():
    from exomedepth_b200 import shard
    for n in (0, 1, 7, 8, 250, 2000):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= -(-n // world) if n else True
