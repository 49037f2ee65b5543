#!/usr/bin/env python
"""bench.py — emission + Viterbi hot path of ExomeDepth on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]             this repo's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...      the reference's own CPU code, all host cores
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ... one rank per GPU (weak scaling)

A "step" is one pass of the hot path over one synthetic cohort batch: per-bin beta-binomial emission
log-likelihood for every copy-number state, then the per-chromosome HMM Viterbi sweep, traceback and
call table with the CallCNVs framing.  Workload at every N: BASELINE.json configs[1] per GPU
(256 samples x 200,000 exon bins x 5 CN states, FP64); samples shard across ranks with no data-path
collective, so scaling is weak.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "exon-bins x samples / s (emission + Viterbi)"
UNIT = "bin*samples/s"
N_SAMPLES, N_BINS, N_STATES = 256, 200_000, 5
TP, CNV_LEN = 1e-4, 50000.0


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU arms
def _cpu_worker(args):
    """One sample through the reference's CPU path exactly as R drives it: one get_loglike_matrix, then one
    C_hmm per chromosome with the CallCNVs framing."""
    kind, s, n_states = args
    from exomedepth_b200 import synth
    from oracle import framing
    g = _cpu_worker.geom
    obs, phi, e = synth.sample(s, g["reference"])
    tot = obs + g["reference"]
    t0 = time.perf_counter()
    if kind == "reference":
        from oracle import ref as impl
        impl.api().quiet(True)
        ll = impl.get_loglike_matrix(np.full(obs.size, phi), np.full(obs.size, e), tot, obs, 1.0)
        T = framing.transition_matrix(TP, 3)
    else:
        from oracle import port as impl
        ll = impl.emission(phi, e, tot, obs, impl.state_odds(n_states))
        T = impl.callcnvs_transitions(n_states, TP)
    ncalls = 0
    off = g["offsets"]
    for c in range(len(off) - 1):
        b0, b1 = off[c], off[c + 1]
        loc, pos = framing.frame_chromosome(ll[b0:b1], g["start"][b0:b1].astype(float), g["end"][b0:b1].astype(float), CNV_LEN)
        ncalls += len(impl.c_hmm(T, loc, pos, CNV_LEN)[1])
    return time.perf_counter() - t0, ncalls


def _cpu_init(n_bins):
    from exomedepth_b200 import synth
    off, start, end = synth.geometry(n_bins)
    _, ref = synth.shared(start.size)
    _cpu_worker.geom = dict(offsets=off, start=start, end=end, reference=ref)


def cpu_throughput(kind, n_states, samples, cores, n_bins=N_BINS, first=0):
    """bin*samples/s of the CPU path over `samples` samples, one sample per worker process."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(n_bins,)) as pool:
        pool.map(_cpu_worker, [(kind, first + i, n_states) for i in range(cores)])        # warm the workers
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(kind, first + i, n_states) for i in range(samples)], chunksize=1)
        wall = time.perf_counter() - t0
    per_core = n_bins / np.mean([r[0] for r in res])
    return n_bins * samples / wall, per_core, wall


def reference_arm(a, rank):
    if rank != 0:
        return
    from oracle import ref
    cores = os.cpu_count() or 1
    kind = "reference" if ref.available() else "port"
    if kind == "reference":
        ref.api()            # dlopen oracle/_ref/libexomedepth_ref.so in THIS process too (the workers are forked from it)
    states = 3 if kind == "reference" else N_STATES
    per_step = max(6 * cores, 16)
    vals, per_core = [], []
    for it in range(a.warmup + a.steps):
        v, pc, wall = cpu_throughput(kind, states, per_step, cores)
        if it >= a.warmup:
            vals.append(v)
            per_core.append(pc)
    value = float(np.mean(vals))
    # the like-for-like denominator of the 5-state GPU arm: the C restatement (oracle port) at 5 states on the same cores —
    # the compiled reference cannot run it (src/hmm.cpp:37-40)
    port5 = None
    if kind == "reference" and not a.no_aux:
        v5, pc5, _ = cpu_throughput("port", N_STATES, per_step, cores)
        port5 = dict(value=v5, unit=UNIT, cores=cores, per_core=pc5, kind="port", states=N_STATES,
                     note="oracle port (plain-C restatement, bit-identical to the compiled reference at 3 states) at the GPU arm's 5 states")
    sample = (f"{per_step} samples x {N_BINS} bins per step, {states} states, one sample per worker process, "
              f"get_loglike_matrix + per-chromosome C_hmm with CallCNVs framing")
    out = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
               ms_per_step=1e3 * per_step * N_BINS / value, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f64", data="synthetic",
               config=dict(workload=f"synthetic {N_SAMPLES} samples x {N_BINS} bins x {N_STATES} CN states per GPU "
                                    f"(BASELINE.json configs[1]); CallCNVs framing, tp=1e-4, L=50000", nproc=cores, reference_states=states,
                           note=("the reference implements 3 states only (src/hmm.cpp:37-40, src/CNV_estimate.cpp:69); it is timed "
                                 "at 3 states on the same synthetic samples, which is LESS work per bin*sample than the 5-state GPU arm"
                                 if kind == "reference" else "compiled reference unavailable: oracle port at 5 states")),
               cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=sample,
                                 per_core=float(np.mean(per_core))),
               e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), port_5_states=port5)
    emit(out)


# ------------------------------------------------------------------------------------------------ GPU arm
def small_panel(dev, ns=512, states=7, steps=50):
    """512 samples x 5,000 bins (chr1-chr5, 1,000 each) x 7 states, device-resident: emission + Viterbi + CallCNVs sums per
    step, as stream launches and as one graph replay."""
    import torch

    import exomedepth_b200 as edb
    from exomedepth_b200 import _lib, synth
    d = synth.cohort(64, per_chrom=(5, 1000))
    reps = ns // 64
    nb = int(d["start"].size)
    nbp = (nb + 15) // 16 * 16
    co = edb.Cohort(d["offsets"], d["start"], d["end"], n_states=states, transition_probability=TP, expected_cnv_length=CNV_LEN)
    obs = torch.from_numpy(np.tile(d["observed"], (reps, 1))).to(dev)
    ref = torch.from_numpy(d["reference"]).to(dev)
    phi, exp = torch.from_numpy(np.tile(d["phi"], reps)).to(dev), torch.from_numpy(np.tile(d["expected"], reps)).to(dev)
    ll = torch.empty((ns, states, nbp), dtype=torch.float64, device=dev)
    path = torch.empty((ns, nbp), dtype=torch.int8, device=dev)
    calls = torch.zeros((ns, 256, 4), dtype=torch.int32, device=dev)
    ncalls = torch.zeros(ns, dtype=torch.int32, device=dev)
    stats = torch.zeros((ns, 256, 3), dtype=torch.float64, device=dev)
    cor = torch.zeros(ns, dtype=torch.float64, device=dev)
    args, kw = (obs, ref, phi, exp, ll, path, calls, ncalls), dict(what=7, call_stats=stats, cor=cor)

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        host_ms = 1e3 * (time.perf_counter() - t0) / steps          # time the host spends enqueueing one step
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, host_ms

    stream_ms, stream_host = timed(lambda: co.run_device(*args, **kw))
    want = (path.clone(), ncalls.clone(), ll.clone())
    _lib.profile(True)
    for _ in range(5):
        co.run_device(*args, **kw)
    kernel_ms = {k: round(v[1] / 5, 5) for k, v in sorted(_lib.profile_read().items(), key=lambda kv: -kv[1][1])}
    _lib.profile(False)
    _lib.launch_count(reset=True)
    gr = co.capture_device(*args, **kw)
    per_step = _lib.launch_count(reset=True)
    graph_ms, graph_host = timed(gr.launch)
    same = bool(torch.equal(path, want[0]) and torch.equal(ncalls, want[1]) and torch.equal(ll.nan_to_num(), want[2].nan_to_num()))
    gr.close()
    co.close()
    cells = ns * nb
    return dict(workload=f"synthetic {ns} samples x {nb} bins x {states} CN states (BASELINE.json configs[3]), device-resident, "
                         "emission + Viterbi + CallCNVs sums", kernels_per_step=int(per_step),
                stream_launch_ms=stream_ms, stream_launch_host_ms=stream_host, graph_replay_ms=graph_ms, graph_replay_host_ms=graph_host,
                stream_launch_value=cells / (stream_ms / 1e3), graph_replay_value=cells / (graph_ms / 1e3), unit=UNIT,
                kernel_ms_per_step=kernel_ms, replay_identical=same, note="CUDA events over 50 back-to-back steps; working set 150 MB > L2")


def check_samples(kind, n_states, d, which, ll, path, calls, ncalls):
    """Oracle comparison of samples `which` of a finished batch (host copies of the GPU outputs): likelihoods within 1e-10,
    Viterbi path and call table equal.  kind = 'reference' (compiled reference, 3 states) or 'port' (C restatement)."""
    from oracle import framing
    if kind == "reference":
        from oracle import ref as impl
        impl.api().quiet(True)
        T = framing.transition_matrix(TP, 3)
    else:
        from oracle import port as impl
        T = impl.callcnvs_transitions(n_states, TP)
    worst, paths_equal, calls_equal, n_cells = 0.0, True, True, 0
    off, start, end, ref = d["offsets"], d["start"], d["end"], d["reference"]
    for k, s in enumerate(which):
        obs = d["observed"][s]
        tot = obs + ref
        if kind == "reference":
            want = impl.get_loglike_matrix(np.full(obs.size, d["phi"][s]), np.full(obs.size, d["expected"][s]), tot, obs, 1.0)
        else:
            want = impl.emission(d["phi"][s], d["expected"][s], tot, obs, impl.state_odds(n_states))
        got = ll[k].T
        assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern"
        ok = np.isfinite(want)
        worst = max(worst, float(np.max(np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1e-2))))
        n_cells += int(ok.sum())
        c = 0
        for ch in range(len(off) - 1):
            b0, b1 = off[ch], off[ch + 1]
            loc, pos = framing.frame_chromosome(got[b0:b1], start[b0:b1].astype(float), end[b0:b1].astype(float), CNV_LEN)
            p, cl = impl.c_hmm(T, loc, pos, CNV_LEN)
            paths_equal &= bool(np.array_equal(path[k, b0:b1], p[1:-1]))
            for (sp, ep, typ, nex) in cl:
                calls_equal &= c < calls.shape[1] and calls[k, c].tolist() == [sp - 1 + b0, ep - 1 + b0, typ, nex]
                c += 1
        calls_equal &= int(ncalls[k]) == c
    return dict(oracle=kind, states=n_states, samples=[int(s) for s in which], cells=n_cells, max_rel=worst,
                within_1e10=bool(worst <= 1e-10), paths_equal=bool(paths_equal), calls_equal=bool(calls_equal))


def gpu_arm(a, rank, world):
    import torch

    import exomedepth_b200 as edb
    from exomedepth_b200 import _lib, synth

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    edb.init(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    S, ns = a.states, a.samples
    # ---- shared exon-bin metadata + reference aggregate: built on rank 0, NCCL-broadcast to the others ------
    from exomedepth_b200 import shard
    arrays = None
    if rank == 0:
        off, start, end = synth.geometry(a.bins)
        arrays = dict(offsets=off, start=start, end=end, reference=synth.shared(start.size)[1])
    shared = shard.broadcast_arrays(arrays, ["offsets", "start", "end", "reference"], dist, device=dev)
    off, start, end, ref = shared["offsets"], shared["start"], shared["end"], shared["reference"]
    nb = int(start.size)
    t0 = time.time()
    # rank 0 builds the host-libm log-transition table; its bytes are broadcast so every rank holds the same bits
    co = shard.make_cohort(shared, dist, device=dev, n_states=S, transition_probability=TP, expected_cnv_length=CNV_LEN)
    table_s = time.time() - t0

    # ---- this rank's samples (weak scaling: `ns` per GPU) -----------------------------------------------------
    obs_h = np.empty((ns, nb), np.int32)
    phi_h, exp_h = np.empty(ns), np.empty(ns)
    for i in range(ns):
        obs_h[i], phi_h[i], exp_h[i] = synth.sample(rank * ns + i, ref)
    obs_t = torch.from_numpy(obs_h).to(dev)
    ref_t = torch.from_numpy(ref).to(dev)
    phi_t, exp_t = torch.from_numpy(phi_h).to(dev), torch.from_numpy(exp_h).to(dev)
    nbp = (nb + 15) // 16 * 16
    ll = torch.empty((ns, S, nbp), dtype=torch.float64, device=dev)
    path = torch.empty((ns, nbp), dtype=torch.int8, device=dev)
    cap = 1024
    calls = torch.zeros((ns, cap, 4), dtype=torch.int32, device=dev)
    ncalls = torch.zeros(ns, dtype=torch.int32, device=dev)

    def step(what=3):
        co.run_device(obs_t, ref_t, phi_t, exp_t, ll, path, calls, ncalls, what=what)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    _lib.launch_count(reset=True)
    _lib.profile(True)                  # one CUDA event before every kernel launch, on the launching stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count()
    prof = _lib.profile_read()          # {kernel: (launches, total ms, longest ms)} over exactly the timed region
    _lib.profile(False)
    # device time per step and kernel.  An unsegmented sweep runs as two CONCURRENT launches per step (longest chromosomes |
    # all others, forked from the same point): the step pays for the longer one, so that is the duration the roofline uses;
    # the segmented sweep (the default at this shape) is one launch.
    kt = {k: (v[2] if k == "viterbi_sweep" and v[0] > a.steps else v[1] / a.steps) for k, v in prof.items()}
    launches_per_step = {k: v[0] / a.steps for k, v in prof.items()}
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank_ms = [ms / a.steps]
    if dist:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(x[0]) / a.steps for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    total_calls = int(ncalls.sum())
    status = _lib.load().edb200_status(0)
    # what the segmented sweep of the last timed step did: pieces, decisions listed with a lead below 2^-14, chains that
    # went to the exact repair pass (include/exomedepth_b200.h: edb200_cohort_segment_stats)
    segments = co.segment_stats()

    # ---- parity of THE BENCHMARKED BUFFERS (untimed): four samples of the batch the timed steps just produced against the C
    # restatement at the benchmark's state count, and the same four samples through the same kernels at 3 states against the
    # compiled reference itself (oracle/_ref, when it travelled to this box)
    parity = None
    if rank == 0 and not a.no_parity:
        from oracle import ref as oref
        which = sorted({0, ns // 3, (2 * ns) // 3, ns - 1})
        dd = dict(offsets=off, start=start, end=end, reference=ref, observed=obs_h, phi=phi_h, expected=exp_h)
        idx = torch.tensor(which, device=dev)
        parity = [check_samples("port", S, dd, which, ll[idx].cpu().numpy()[:, :, :nb], path[idx].cpu().numpy()[:, :nb],
                                calls[idx].cpu().numpy(), ncalls[idx].cpu().numpy())]
        if oref.available():
            co3 = edb.Cohort(off, start, end, n_states=3, transition_probability=TP, expected_cnv_length=CNV_LEN)
            k = len(which)
            ll3 = torch.empty((k, 3, nbp), dtype=torch.float64, device=dev)
            path3 = torch.empty((k, nbp), dtype=torch.int8, device=dev)
            calls3 = torch.zeros((k, cap, 4), dtype=torch.int32, device=dev)
            ncalls3 = torch.zeros(k, dtype=torch.int32, device=dev)
            co3.run_device(obs_t[idx].contiguous(), ref_t, phi_t[idx].contiguous(), exp_t[idx].contiguous(), ll3, path3, calls3, ncalls3,
                           what=3, mode=_lib.EMISSION_TABLE)
            torch.cuda.synchronize()
            parity.append(check_samples("reference", 3, dd, which, ll3.cpu().numpy()[:, :, :nb], path3.cpu().numpy()[:, :nb],
                                        calls3.cpu().numpy(), ncalls3.cpu().numpy()))
            co3.close()

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D of the counts inside the timed region, results back
    # in host memory.  Primary = what CallCNVs returns to the user (R/class_definition.R:311-419): the CNV call table with
    # its BF / reads.* columns and cor(test, reference) — those sums are taken on the device (callcnvs.cu), so the FP64
    # likelihood matrix stays in HBM as the cohort's resident `likelihood` slot.  Variants: + the per-bin Viterbi path
    # (C_hmm's first return value), + the likelihood matrix itself (get_loglike_matrix's return value, 8*S bytes per
    # bin*sample over PCIe: that copy alone is ~37 ms).
    e2e = None
    if not a.no_e2e:
        hb = _lib.PinnedPool()
        obs_p = hb.empty((ns, nb), np.int32)
        obs_p[:] = obs_h
        # the ingestion layouts, written once per cohort by the loader: rows of 12-bit fields + overflow list (edb200_batch.observed12,
        # the default here) and 16-bit counts + overflow list (observed16)
        obs16_p, ovf_i, ovf_v = edb.pack_counts(obs_h, out=hb.empty((ns, nb), np.uint16))
        base = dict(calls=hb.empty((ns, cap, 4), np.int32), ncalls=hb.empty((ns,), np.int32),
                    call_stats=hb.empty((ns, cap, 3), np.float64), cor=hb.empty((ns,), np.float64))
        n_e2e = max(2, min(a.steps, 10))

        pairs12 = (nb + 1) // 2
        out12 = hb.empty((ns, (3 * pairs12 + 3) // 4 * 4), np.uint8)
        obs12_p, ovf12_i, ovf12_v = edb.pack_counts12(obs_h, out=out12)
        t0 = time.perf_counter()
        edb.pack_counts12(obs_h, out=out12)
        pack_ms = 1e3 * (time.perf_counter() - t0)

        h2d = obs12_p.nbytes + ovf12_i.nbytes + ovf12_v.nbytes + ref.nbytes + phi_h.nbytes + exp_h.nbytes

        def timed(out, counts=None, **kw):
            counts = obs12_p if counts is None else counts
            ovf = (ovf_i, ovf_v) if counts is obs16_p else (ovf12_i, ovf12_v) if counts is obs12_p else None
            co.run_host(counts, ref, phi_h, exp_h, call_cap=cap, out=out, want_stats=True, overflow=ovf, **kw)     # warm-up (allocations)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                co.run_host(counts, ref, phi_h, exp_h, call_cap=cap, out=out, want_stats=True, overflow=ovf, **kw)
            barrier()
            t = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
            if dist:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            assert int(out["ncalls"].sum()) == total_calls
            return dict(value=world * ns * nb / float(t[0]), unit=UNIT, ms_per_step=1e3 * float(t[0]),
                        d2h_bytes_per_step=int(sum(v.nbytes for v in out.values())))

        e2e = timed(base, want_ll=False, want_path=False)
        e2e.update(h2d_bytes_per_step=int(h2d), steps=n_e2e,
                   pack_ms_once_per_cohort=pack_ms, pack_note="edb200_pack_counts12 on the host's threads: int32 count matrix -> the 12-bit layout; run "
                                                              "once per cohort by the loader, NOT inside the timed call (e2e.int32_counts is the call on the "
                                                              "unpacked matrix)",
                   counts_layout=f"rows of 12-bit fields [sample][bin] + overflow list ({int(ovf12_i.size)} entries of 4095 reads and more) — "
                                 "edb200_batch.observed12",
                   api="edb200_cohort_run_host (C ABI, pinned host buffers): counts in; CNV call table, per-call BF / reads.expected / "
                       "reads.observed sums and cor(test, reference) out — the output of CallCNVs; likelihood matrix resident in HBM; "
                       "sample chunks pipelined over PCIe (upload k+1 | emission k+1 | segmented Viterbi k)",
                   segmented_sweep=co.segment_stats())
        e2e["int32_counts"] = dict(timed(base, counts=obs_p, want_ll=False, want_path=False), h2d_bytes_per_step=int(obs_p.nbytes + ref.nbytes + phi_h.nbytes + exp_h.nbytes),
                                   note="the same call with the counts as int32 [sample][bin] (edb200_batch.observed)")
        e2e["uint16_counts"] = dict(timed(base, counts=obs16_p, want_ll=False, want_path=False),
                                    h2d_bytes_per_step=int(obs16_p.nbytes + ovf_i.nbytes + ovf_v.nbytes + ref.nbytes + phi_h.nbytes + exp_h.nbytes),
                                    note=f"the same call with the counts as uint16 [sample][bin] + overflow list ({int(ovf_i.size)} entries) — "
                                         "edb200_batch.observed16: a third more bytes over PCIe")
        with_path = dict(base, path=hb.empty((ns, nb), np.int8))
        e2e["with_path"] = dict(timed(with_path, want_ll=False, want_path=True), note="+ per-bin Viterbi path (int8) copied back")
        if not a.no_ll:
            with_ll = dict(with_path, ll=hb.empty((ns, S, nb), np.float64))
            e2e["with_ll_copy"] = dict(timed(with_ll, want_ll=True, want_path=True),
                                       note="+ FP64 likelihood matrix copied back (8*S bytes per bin*sample: bound by the PCIe device-to-host copy)")
        hb.close()

    # ---- the rows around the hot path, device-resident on this rank's cohort: reported, not part of `value`
    aux = None
    if rank == 0 and not a.no_aux:
        from exomedepth_b200 import refset
        bl = (end - start + 1).astype(np.float64)
        sel = refset.select_bins(obs_h.sum(0, dtype=np.int64), bl)
        sel_t, bl_t = torch.from_numpy(sel).to(dev), torch.from_numpy(bl).to(dev)
        kp = refset.kpad(sel.size)
        z = torch.empty((ns, kp), dtype=torch.float64, device=dev)
        cmat = torch.empty((ns, ns), dtype=torch.float64, device=dev)
        mu_t, phi2_t, ll_t = (torch.empty(ns, dtype=torch.float64, device=dev) for _ in range(3))
        info_t = torch.empty(ns, dtype=torch.int32, device=dev)
        L = _lib.load()
        st0 = torch.cuda.current_stream().cuda_stream

        def aux_step():
            _lib.check(L.edb200_betabin_fit_device(obs_t.data_ptr(), obs_t.stride(0), ref_t.data_ptr(), 0, ns, nb, mu_t.data_ptr(),
                                                   phi2_t.data_ptr(), ll_t.data_ptr(), info_t.data_ptr(), st0), "betabin_fit")
            refset.standardize_device(obs_t, sel_t, bl_t, z)
            refset.gram_device(z, z, sel.size, cmat)

        aux_step()
        torch.cuda.synchronize()
        _lib.profile(True)
        for _ in range(3):
            aux_step()
        pa = {k: v[1] / v[0] for k, v in _lib.profile_read().items()}
        _lib.profile(False)
        aux = dict(betabin_fit_ms=pa["betabin_fit"], betabin_fit_iterations_max=int(info_t.max()),
                   refset_standardize_ms=pa["refset_standardize"], refset_gram_ms=pa["refset_gram"] + pa["refset_reduce"],
                   refset_gram_tflops_fp64=2.0 * ns * ns * kp / pa["refset_gram"] / 1e9, refset_selected_bins=int(sel.size),
                   note="beta-binomial fit (aod::betabin stand-in) and select.reference.set correlation sweep of the same cohort; "
                        "device-resident, CUDA events, not part of `value`")

        # ---- the drop-in shape R gets from r_glue.c: ONE sample per .Call — get_loglike_matrix, then C_hmm per chromosome
        # (R/class_definition.R:184-189, R/tools.R:97), host buffers in and out, every call its own round trip
        try:
            from oracle import framing
            n_ps = 5
            per = []
            for s_i in range(n_ps):
                t0 = time.perf_counter()
                ll_h = edb.get_loglike_matrix(np.full(nb, phi_h[s_i]), np.full(nb, exp_h[s_i]), obs_h[s_i] + ref, obs_h[s_i], 1.0)
                T3 = framing.transition_matrix(TP, 3)
                for ch in range(len(off) - 1):
                    b0, b1 = off[ch], off[ch + 1]
                    loc, pos = framing.frame_chromosome(ll_h[b0:b1], start[b0:b1].astype(float), end[b0:b1].astype(float), CNV_LEN)
                    edb.C_hmm(3, loc.shape[0], T3, loc, pos, CNV_LEN)
                per.append(time.perf_counter() - t0)
            dt = float(np.mean(per[1:]))
            aux["per_sample_call_shape"] = dict(ms_per_sample=1e3 * dt, first_sample_ms=1e3 * per[0], value=nb / dt, unit=UNIT, states=3,
                                                calls_per_sample=1 + len(off) - 1,
                                                note="edb200_get_loglike_matrix + one edb200_hmm per chromosome per sample (the two-routine drop-in "
                                                     "of src/ExomeDepth_init.c:14-24), incl. the Python framing of R/class_definition.R:364-368; "
                                                     "the first sample builds the log-transition table of every chromosome (kept per positions / "
                                                     "matrix / length: the later samples of a run reuse it), ms_per_sample is the mean of the others")
        except Exception as e:                                          # noqa: BLE001
            aux["per_sample_call_shape"] = dict(error=f"{type(e).__name__}: {e}")

        # ---- reference-API-faithful cohort: per-bin phi / expected vectors (SURVEY.md §8d "+20 B" variant), 64 samples
        try:
            npb = min(ns, 64)
            phi_pb = phi_t[:npb, None].expand(npb, nbp).contiguous()
            exp_pb = exp_t[:npb, None].expand(npb, nbp).contiguous()
            bpb = _lib.Batch(npb, obs_t.data_ptr(), obs_t.stride(0), ref_t.data_ptr(), 0, phi_pb.data_ptr(), exp_pb.data_ptr(), ll.data_ptr(),
                             ll.stride(1), path.data_ptr(), path.stride(0), calls.data_ptr(), ncalls.data_ptr(), cap, None, None, nbp)
            import ctypes as C
            run_pb = lambda: _lib.check(L.edb200_cohort_run_device(co.handle, C.byref(bpb), 3, 0, st0), "per-bin run")
            run_pb()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(3):
                run_pb()
            ev1.record()
            torch.cuda.synchronize()
            ms_pb = ev0.elapsed_time(ev1) / 3
            aux["per_bin_phi_expected"] = dict(samples=npb, ms_per_step=ms_pb, value=npb * nb / (ms_pb / 1e3), unit=UNIT,
                                               bytes_per_unit=4 + 8 * S + 1 + 8 * S + 16,
                                               note="phi / expected as per-bin vectors per sample (edb200_batch.per_bin_stride): per-state constants "
                                                    "rebuilt per bin in registers — FP64-pipe bound, not HBM bound")
            step()                                                      # leave the scalar results in the output buffers
        except Exception as e:                                          # noqa: BLE001
            aux["per_bin_phi_expected"] = dict(error=f"{type(e).__name__}: {e}")

        # ---- the north star's single-GPU configuration: 2,000 samples x 200k bins x 5 states on ONE B200 (the 256 samples tiled)
        try:
            if S == N_STATES and ns == N_SAMPLES and not a.no_large:
                big = 2000
                rep = -(-big // ns)
                obs_b = obs_t.repeat(rep, 1)[:big].contiguous()
                phi_b, exp_b = phi_t.repeat(rep)[:big].contiguous(), exp_t.repeat(rep)[:big].contiguous()
                ll_b = torch.empty((big, S, nbp), dtype=torch.float64, device=dev)
                path_b = torch.empty((big, nbp), dtype=torch.int8, device=dev)
                calls_b = torch.zeros((big, cap, 4), dtype=torch.int32, device=dev)
                ncalls_b = torch.zeros(big, dtype=torch.int32, device=dev)
                run_b = lambda: co.run_device(obs_b, ref_t, phi_b, exp_b, ll_b, path_b, calls_b, ncalls_b, what=3)
                run_b()
                torch.cuda.synchronize()
                _lib.profile(True)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for _ in range(3):
                    run_b()
                ev1.record()
                torch.cuda.synchronize()
                pk = {k: round(v[1] / 3, 4) for k, v in _lib.profile_read().items()}
                _lib.profile(False)
                ms_b = ev0.elapsed_time(ev1) / 3
                peak_b, _src = peaks()
                aux["large_cohort_one_gpu"] = dict(
                    workload=f"{big} samples x {nb} bins x {S} states on one GPU (north star; the {ns} synthetic samples tiled), device-resident",
                    ms_per_step=ms_b, value=big * nb / (ms_b / 1e3), unit=UNIT, kernel_ms_per_step=pk,
                    sweep_hbm_frac=big * nb * (8 * S + 1) / (pk.get("viterbi_sweep", float("nan")) / 1e3) / 1e9 / peak_b,
                    emission_hbm_frac=big * nb * (4 + 8 * S) / (pk.get("emission", float("nan")) / 1e3) / 1e9 / peak_b,
                    calls_match_tiled=bool(int(ncalls_b.sum()) == int(sum(int(ncalls[i % ns]) for i in range(big)))))
                del obs_b, ll_b, path_b, calls_b
        except Exception as e:                                          # noqa: BLE001
            aux["large_cohort_one_gpu"] = dict(error=f"{type(e).__name__}: {e}")

        # BASELINE.json configs[3], the launch-bound regime: plain stream launches against the replay of a captured CUDA
        # graph (edb200_cohort_capture_device).  Reported only; a failure here must not take the bench line with it.
        try:
            aux["small_panel"] = small_panel(dev)
        except Exception as e:                                          # noqa: BLE001
            aux["small_panel"] = dict(error=f"{type(e).__name__}: {e}")

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    cells = ns * nb
    value = world * cells * a.steps / (ms / 1e3)
    peak, peak_src = peaks()
    # algorithmic bytes per bin*sample: emission reads the test count and writes S log-likelihoods;
    # the Viterbi sweep reads them back and the path byte leaves downstream of it
    alg = {"emission": 4 + 8 * S, "viterbi_sweep": 8 * S + 1}
    dom = max(alg, key=lambda k: kt.get(k, 0.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass

    def roof(name):
        ach = cells * alg[name] / (kt[name] / 1e3) / 1e9
        return dict(kernel=name, bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                    traffic=(traffic or {}).get(name) if isinstance(traffic, dict) else None, ms_per_launch=kt[name],
                    launches_per_step=launches_per_step[name], bytes_per_unit=alg[name], peak_source=peak_src,
                    note="achieved = algorithmic bytes per step / this kernel's device time per step (CUDA events on its stream; "
                         "concurrent launches of one step: the longest); traffic = ncu dram bytes per step (profiles/traffic.json)")

    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup, ms_per_step=ms / a.steps,
               higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
               config=dict(workload=f"synthetic {ns} samples x {nb} bins x {S} CN states per GPU (BASELINE.json configs[1]); "
                                    "CallCNVs framing, tp=1e-4, L=50000",
                           cache="inputs+outputs per step (2.3 GB) exceed the 126 MB L2; no flush needed",
                           shared_metadata="bin geometry, reference aggregate and host-libm log-transition table NCCL-broadcast from rank 0"
                           if world > 1 else "single rank", table_build_s=table_s, total_calls=total_calls, status=status, nproc=os.cpu_count()),
               clocks=clocks, e2e=e2e, gpu_launches=int(launches), parity=parity, segmented_sweep=segments, roofline=roof(dom),
               roofline_other=roof("emission" if dom != "emission" else "viterbi_sweep"),
               kernel_ms_per_step={k: round(v, 5) for k, v in sorted(kt.items(), key=lambda kv: -kv[1])},
               ms_per_step_per_rank=[round(v, 4) for v in per_rank_ms], aux=aux)
    if world == 1 and not a.no_cpu:
        from oracle import ref as oref
        cores = os.cpu_count() or 1
        kind = "reference" if oref.available() else "port"
        if kind == "reference":
            oref.api()       # the library is then mapped in this process as well as in the forked workers
        states = 3 if kind == "reference" else S
        nsamp = max(6 * cores, 16)
        v, pc, wall = cpu_throughput(kind, states, nsamp, cores)
        out["cpu_baseline"] = dict(value=v, unit=UNIT, cores=cores, kind=kind, per_core=pc, wall_s=wall,
                                   sample=f"{nsamp} of the same synthetic samples x {nb} bins, {states} states "
                                          f"(the reference implements 3 states only), one sample per worker process")
    emit(out)
    if dist:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ other workloads
def _dist_setup(world):
    import torch
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    return local, dev, dist


def fp64_peak():
    """Measured FP64 FMA issue rate of one B200 SM (tools/ubench/fp64tput.cu, profiles/r2b_fp64tput.txt): 1.845 warp
    instructions per clock and SM -> TFLOP/s at the SM clock the run saw."""
    return 1.845 * 32 * 2 * 148 * 1.965e9 / 1e12, "measured DFMA issue rate, profiles/r2b_fp64tput.txt (1.845 warp-instr/clk/SM x 148 SMs x 1.965 GHz)"


def refset_arm(a, rank, world):
    """BASELINE.json configs[4]: the select.reference.set correlation sweep (R/optimize_reference_set.R:100) of a
    2,000-sample x 200k-bin cohort, leave-one-out over the whole cohort, samples sharded over the ranks (strong scaling:
    the cohort is fixed).  A step = standardise this rank's rows, exchange, form this rank's block of the N x N matrix."""
    import torch

    import exomedepth_b200 as edb
    from exomedepth_b200 import _lib, refset, shard, synth
    local, dev, dist = _dist_setup(world)
    edb.init(local)
    n_total, nb = a.samples if a.samples != N_SAMPLES else 2000, a.bins
    d = synth.cohort(16, n_bins=nb)
    nb = int(d["start"].size)
    lo, hi = shard.shard_range(n_total, rank, world)
    rng = np.random.default_rng(7)
    thin = rng.uniform(0.6, 1.0, n_total)
    counts = np.empty((hi - lo, nb), np.int32)
    for s in range(lo, hi):                                 # every rank makes only its block; thinning keeps samples distinct
        counts[s - lo] = np.random.default_rng(1000 + s).binomial(d["observed"][s % 16], thin[s])
    bl = (d["end"] - d["start"] + 1).astype(np.float64)
    total = torch.from_numpy(counts.sum(0, dtype=np.int64)).to(dev)
    if dist:
        dist.all_reduce(total)                              # per-bin totals of the whole cohort (the bin filter needs them)
    sel = refset.select_bins(total.cpu().numpy(), bl)
    per = -(-n_total // world)
    kp = refset.kpad(sel.size)
    n_local = hi - lo
    c_t, sel_t, bl_t = torch.from_numpy(counts).to(dev), torch.from_numpy(sel).to(dev), torch.from_numpy(bl).to(dev)
    z_local = torch.zeros((per, kp), dtype=torch.float64, device=dev)
    z_all = torch.empty((world * per, kp), dtype=torch.float64, device=dev) if dist else z_local
    out = torch.empty((max(n_local, 1), n_total), dtype=torch.float64, device=dev)
    fused = dist is not None and per <= 256 and world <= 4 and not a.no_fused and a.exchange in ("auto", "fused")
    by_blocks = dist is not None and not fused and (a.exchange == "blocks" or (a.exchange == "auto" and world > 2))
    blk_bufs = blk_outs = None
    if by_blocks:
        rows = [max(0, min(per, n_total - j * per)) for j in range(world)]
        blk_bufs = [z_local if j == rank else torch.empty_like(z_local) for j in range(world)]
        blk_outs = [torch.empty((n_local, rows[j]), dtype=torch.float64, device=dev) for j in range(world)]
    if fused:
        z_ptr, handle = refset.block_alloc(per, sel.size)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        refset.peers_open(handles, rank)

    def step():
        if fused:
            refset.standardize_device(c_t, sel_t, bl_t, z_ptr)
            torch.cuda.synchronize()
            dist.barrier()                                  # every block complete before any rank reads it over NVLink
            refset.gram_peers_device(n_local, per, n_total, sel.size, out)
        elif by_blocks:
            refset.standardize_device(c_t, sel_t, bl_t, z_local[:n_local])
            shard.exchange_and_gram(z_local, n_local, per, n_total, sel.size, dist, dev, bufs=blk_bufs, outs=blk_outs)
        else:
            refset.standardize_device(c_t, sel_t, bl_t, z_local[:n_local])
            if dist:
                dist.all_gather_into_tensor(z_all, z_local)
            refset.gram_device(z_local[:n_local], z_all[:n_total], sel.size, out)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    _lib.launch_count(reset=True)
    _lib.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count()
    prof = _lib.profile_read()
    _lib.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / a.steps
    # end to end through the host-pointer entry point (counts in pinned host memory -> correlation rows back), rank 0's block
    e2e = None
    if not a.no_e2e and rank == 0:
        n_e = min(n_local, 64)
        hb = _lib.PinnedPool()
        cp = hb.empty((n_total if world == 1 else n_local, nb), np.int32)
        cp[:] = counts if world > 1 else counts
        t0 = time.perf_counter()
        refset.correlations(cp, sel, bl, row0=0, n_rows=n_e)
        dt = time.perf_counter() - t0
        e2e = dict(value=n_e * cp.shape[0] * sel.size / dt, unit="pair*bins/s", h2d_bytes_per_step=int(cp.nbytes + sel.nbytes + bl.nbytes),
                   d2h_bytes_per_step=int(n_e * cp.shape[0] * 8), ms_per_step=1e3 * dt,
                   api=f"edb200_refset_correlations (C ABI, pinned host counts): {n_e} test rows against {cp.shape[0]} samples of this rank")
        hb.close()
    if fused:
        barrier()
        refset.peers_close()
    if rank == 0:
        pairs = float(n_total) * n_total
        flops = 2.0 * pairs * kp
        gram_ms = prof["refset_gram"][1] / a.steps
        peak, src = fp64_peak()
        ach = 2.0 * n_local * n_total * kp / (gram_ms / 1e3) / 1e12
        out_line = dict(metric="select.reference.set correlation sweep: sample pairs x selected bins / s", value=pairs * sel.size / (ms / 1e3),
                        unit="pair*bins/s", n_gpus=world, steps=a.steps, warmup=a.warmup, ms_per_step=ms, higher_is_better=True, scaling="strong",
                        vs_baseline=None, dtype="f64", data="synthetic",
                        config=dict(workload=f"select.reference.set sweep, {n_total} samples x {nb} bins ({sel.size} selected), leave-one-out over the "
                                             "cohort (BASELINE.json configs[4])", exchange="CUDA-IPC peer-memory Gram (no all-gather)" if fused else
                                             "one NCCL broadcast per rank's block, block j's Gram while block j + 1 is in flight" if by_blocks else
                                             ("NCCL all-gather of standardised rows, then Gram" if dist else "single rank"),
                                    all_gather_bytes_per_rank=int((world - 1) * per * kp * 8), cache="Z (2.2 GB) exceeds the 126 MB L2", nproc=os.cpu_count()),
                        clocks=clocks, e2e=e2e, gpu_launches=int(launches),
                        roofline=dict(kernel="refset_gram", bound="fp64", achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak, traffic=None,
                                      ms_per_launch=gram_ms, flops_per_step_all_ranks=flops, peak_source=src,
                                      note="FP64 tensor-core contraction (mma.sync m8n8k4; there is no FP64 tcgen05 kind); achieved = this rank's 2*m*n*k flops / its Gram kernel time"),
                        kernel_ms_per_step={k: round(v[1] / a.steps, 5) for k, v in prof.items()})
        if world == 1 and not a.no_cpu:
            from oracle import refset as oref
            m = 48
            t0 = time.perf_counter()
            oref.cohort_correlations(counts[:m], bl)
            dt = time.perf_counter() - t0
            out_line["cpu_baseline"] = dict(value=m * m * sel.size / dt, unit="pair*bins/s", cores=1, kind="port", wall_s=dt,
                                            sample=f"{m} of the same samples, numpy restatement of R/optimize_reference_set.R:81-100 (oracle/refset.py)")
        emit(out_line)
    if dist:
        dist.destroy_process_group()


def small_panel_arm(a, rank, world):
    """BASELINE.json configs[3]: 512 samples x 5,000 bins (chr1-chr5, 1,000 each) x 7 states per GPU, device-resident,
    emission + Viterbi + CallCNVs sums per step, replayed from a captured CUDA graph (the launch-bound regime)."""
    import torch

    import exomedepth_b200 as edb
    local, dev, dist = _dist_setup(world)
    edb.init(local)
    ns = a.samples if a.samples != N_SAMPLES else 512
    r = small_panel(dev, ns=ns, states=7, steps=max(a.steps, 20))
    t = torch.tensor([r["graph_replay_ms"]], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        nb, S = 5000, 7
        cells = ns * nb
        peak, src = peaks()
        kt = r["kernel_ms_per_step"]
        alg = {"emission_panel": 4 + 8 * S, "viterbi_sweep": 8 * S + 1}
        dom = max(alg, key=lambda k: kt.get(k, 0.0))
        ach = cells * alg[dom] / (kt[dom] / 1e3) / 1e9
        line = dict(metric=METRIC, value=world * cells / (ms / 1e3), unit=UNIT, n_gpus=world, steps=max(a.steps, 20), warmup=5, ms_per_step=ms,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=r["workload"], launch="one CUDA-graph replay per step", cache=r["note"], nproc=os.cpu_count()),
                    gpu_launches=int(r["kernels_per_step"] * max(a.steps, 20)), e2e=None,
                    roofline=dict(kernel=dom, bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=None,
                                  ms_per_launch=kt[dom], bytes_per_unit=alg[dom], peak_source=src),
                    kernel_ms_per_step=kt, stream_launch_ms=r["stream_launch_ms"], host_ms_per_step=dict(stream=r["stream_launch_host_ms"], graph=r["graph_replay_host_ms"]),
                    replay_identical=r["replay_identical"])
        if world == 1 and not a.no_cpu:
            from oracle import ref as oref
            cores = os.cpu_count() or 1
            kind = "reference" if oref.available() else "port"
            if kind == "reference":
                oref.api()
            v, pc, wall = cpu_throughput(kind, 3 if kind == "reference" else 7, max(8 * cores, 64), cores, n_bins=nb)
            line["cpu_baseline"] = dict(value=v, unit=UNIT, cores=cores, kind=kind, per_core=pc, wall_s=wall,
                                        sample=f"{max(8 * cores, 64)} synthetic samples x {nb} bins, 3 states (the reference implements 3 only)")
        emit(line)
    if dist:
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """Libraries write banners to file descriptor 1 (NCCL prints its version there at NCCL_DEBUG=VERSION, which this
    image sets): everything but the final JSON line is sent to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(out):
    print(json.dumps(out), file=_REAL_STDOUT or sys.stdout, flush=True)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--samples", type=int, default=N_SAMPLES, help="samples per GPU")
    ap.add_argument("--bins", type=int, default=N_BINS)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the benchmarked buffers")
    ap.add_argument("--no-aux", action="store_true", help="skip the timings of the fit / reference-set kernels")
    ap.add_argument("--no-ll", action="store_true", help="skip the e2e variant that copies the likelihood matrix back")
    ap.add_argument("--no-large", action="store_true", help="skip the 2,000-sample single-GPU aux measurement")
    ap.add_argument("--workload", default="cohort", choices=["cohort", "small_panel", "refset"],
                    help="cohort = BASELINE.json configs[1] per GPU (the metric's configuration; the default); small_panel = configs[3]; "
                         "refset = configs[4], the select.reference.set correlation sweep of 2,000 samples sharded over the ranks")
    ap.add_argument("--states", type=int, default=N_STATES, help="copy-number states of the cohort workload (3 = the reference's own model)")
    ap.add_argument("--no-fused", action="store_true", help="refset: NCCL all-gather + Gram instead of the peer-memory Gram")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "fused", "blocks"],
                    help="refset: how the standardised rows reach the other ranks (auto: fused up to 4 ranks, block-wise broadcasts beyond)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    a.warmup = max(a.warmup, 3) if a.impl == "graft" else a.warmup
    if a.impl == "reference":
        reference_arm(a, rank)
    elif a.workload == "refset":
        refset_arm(a, rank, world)
    elif a.workload == "small_panel":
        small_panel_arm(a, rank, world)
    else:
        gpu_arm(a, rank, world)


if __name__ == "__main__":
    main()
