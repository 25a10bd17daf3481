#!/usr/bin/env python
"""
bench.py — Kalman-filter throughput of the B200 LGSSM path on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE configs[1] — GP(Matern52) (D = 3, scalar observations),
RegularSpacing(0, 0.01, T = 10^7), sigma^2 = 0.1, FP64, synthetic y (circulant-embedding draw of the
same GP + noise). One "step" = one logpdf(model, y) call = T Kalman filter steps (predict + update +
log marginal likelihood), i.e. the reference's headline benchmark (README "logpdf", N = 10^7).

  value      steps/s with y resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e        the same through the public API with a pinned HOST y: H2D copy + kernels + D2H of lml
  roofline   dominant kernel: algorithmic bytes / its mean device time, vs MEASURED_PEAKS.json HBM
  cpu_baseline  the C restatement of the reference's sequential SArrayStorage path (oracle/), 1 thread
  filter_emit   (extra) tgp_filter emitting (m_f, P_f): 104 B/step algorithmic

N > 1 (torchrun): ONE series of N*T steps sharded over time, one rank per GPU ("weak" scaling: T per GPU fixed = `value`), and
BASELINE config 4 beside it on every line (`strong_scaling`: the same 8e7-step series at every N). One kernel launch per shard and
step; the <= 3072 observations that precede a shard travel over NVLink inside the kernels (tgp_fir.cuh).
Both series are checked against the sequential oracle at every N (`lml_rel_err_vs_oracle`).
--impl reference: the oracle port of the reference's CPU path on the host (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_DEFAULT = 10_000_000
DT, SIGMA2 = 0.01, 0.1
N_BUF = 4  # rotating input buffers: 4 x 80 MB = 320 MB > 126 MB L2


def synth_y(T, seed):
    """Synthetic observations: stationary Matern-5/2 draw by circulant embedding (+ noise)."""
    rng = np.random.default_rng(seed)
    n = T + (T & 1)
    w = 2.0 * np.pi * np.fft.rfftfreq(n, d=DT)
    s = (5.0 + w * w) ** -3.0
    z = rng.standard_normal(len(w)) + 1j * rng.standard_normal(len(w))
    f = np.fft.irfft(np.sqrt(s) * z, n=n)[:T]
    f *= 1.0 / f.std()
    return f + math.sqrt(SIGMA2) * rng.standard_normal(T)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        smax = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json)"
    return 6650.0, "of fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_reference(T, reps, seed=20261017 + 2):
    """The reference's sequential logpdf (scan.jl:15-28 + lgssm.jl:147-165) as restated in oracle/ (C, fixed-size
    D = 3 instantiation = SArrayStorage analogue), single-threaded like the reference."""
    from oracle import c_oracle, tgp_oracle as O
    mo = O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, DT, T), SIGMA2)
    cm = c_oracle.Model.from_lgssm(mo)
    y = synth_y(T, seed)
    c_oracle.logpdf(cm, y[: min(T, 100_000)])  # warm
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        lml = c_oracle.logpdf(cm, y)
        ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)), lml


METRIC = "Kalman filter steps/s (logpdf, Matern52 d=3, T=1e7 per GPU, FP64)"


def workload_config(T, world, algo="auto"):
    """`config` of the JSON line — identical in both arms so the driver can pair them."""
    Tglob = T * world
    return {"workload": (f"cfg2: GP(Matern52) D=3 M=1 RegularSpacing(0,{DT},T) sigma2={SIGMA2} logpdf; "
                         f"T={T} per GPU, one series of {Tglob} steps sharded over time") if world > 1 else
                        f"cfg2: GP(Matern52) D=3 M=1 RegularSpacing(0,{DT},{T}) sigma2={SIGMA2} logpdf",
            "T_per_gpu": T, "T_total": Tglob, "algo": algo,
            "l2": f"{N_BUF} rotating input buffers of {8 * T / 1e6:.0f} MB (> 126 MB L2 between reuses)",
            "parallelism": f"time-sharded x{world}" if world > 1 else "single GPU"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T = args.T
    for _ in range(max(args.warmup, 0)):
        cpu_reference(min(T, 1_000_000), 1)
    t, _ = cpu_reference(T, args.steps)
    v = T / t
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(T, max(int(os.environ.get("WORLD_SIZE", "1")), 1), args.algo),
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": 1, "kind": "port",
                         "sample": f"{args.steps} x full T={T} logpdf, C restatement (oracle/lgssm_ref.c, static D=3), 1 thread: the "
                                   "reference recursion is sequential and single-threaded; Julia itself is not installable here"},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"nproc": os.cpu_count()},
    }))


# ------------------------------------------------------------------------------------------------
STRONG_T = 80_000_000      # BASELINE config 4: one series of 8e7 steps, time-sharded over the GPUs (fixed total = strong scaling)
STRONG_CHUNK = 10_000_000


def strong_chunk(c, buf):
    """Chunk c (1e7 steps) of the config-4 series, buffer set `buf`: the same for every world size."""
    return synth_y(STRONG_CHUNK, 20261017 + 4 + 131 * c + 7919 * buf)


def cpu_all_cores(T, nthreads, seed=20261017 + 2):
    """`nthreads` independent logpdf evaluations, one host thread each (oracle/lgssm_ref.c: the only way the single-threaded
    reference can use all cores for this path; labelled "replicas" — NOT what it does for one series)."""
    from oracle import c_oracle, tgp_oracle as O
    cm = c_oracle.Model.from_lgssm(O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, DT, T), SIGMA2))
    ys = np.stack([synth_y(T, seed + 17 * k) for k in range(nthreads)])
    c_oracle.logpdf_replicas(cm, ys[:, : T])      # warm
    t0 = time.perf_counter()
    _, n = c_oracle.logpdf_replicas(cm, ys)
    dt = time.perf_counter() - t0
    return nthreads * T / dt, n


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g

    pkg = g.load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    T = args.T
    K, W = args.steps, max(args.warmup, 3)
    h = pkg.default_handle(local)
    # One real stream for the library's kernels, torch's copies / collectives and the timing events (torch's default stream has
    # handle 0, which the library would read as "use the handle's own stream").
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    if args.algo == "scan":
        h.set_algo(pkg.TGP_ALGO_SCAN)
    f = pkg.to_sde(pkg.GP(pkg.Matern52Kernel()), pkg.B200Storage(local))
    from temporalgps_jl_b200 import sharded

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step, finish):
        """W warm-up steps, then exactly K steps between two events on the launching stream, barrier + synchronize on both sides;
        `finish` (inside the timed region) makes the last step's result available. -> ms per step, max over ranks."""
        for i in range(W):
            step(i)
        finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(K):
            step(i)
        finish()
        e1.record(stream)
        barrier()
        h.synchronize()
        return max_over_ranks(e0.elapsed_time(e1)) / K

    # ================= weak series (the headline `value`): config 2 per GPU, T = 1e7 each, ONE series of world * T steps ==========
    Tglob = T * world
    fx = f(pkg.RegularSpacing(0.0, DT, T), SIGMA2)
    mm = pkg.lgssm._Marshalled(fx.build_lgssm())
    seed_of = lambda r, i: 20261017 + 2 + 97 * i + 1000 * r      # noqa: E731
    ys_host = [synth_y(T, seed_of(rank, i)) for i in range(N_BUF)]
    HALO = sharded.TGP_SHARD_HALO
    # Overlapped scatter (N > 1): every rank's buffer is [the 3072 observations before its shard | its shard]; the kernels read the
    # halo in place, so no GPU ever waits for another. The views below are the shards themselves (what the exchange layout uses too).
    halos = [synth_y(T, seed_of(rank - 1, i))[-HALO:] if rank > 0 else np.zeros(HALO) for i in range(N_BUF)] if world > 1 else None
    ys_pad = [torch.from_numpy(np.concatenate([halos[i], v]) if world > 1 else v).to(dev) for i, v in enumerate(ys_host)]
    ys_dev = [p[HALO:] if world > 1 else p for p in ys_pad]
    lml_dev = torch.zeros(1, dtype=torch.float64, device=dev)
    lml_host = np.zeros(1)
    sh = None
    if world == 1:
        # device-resident inputs AND output: each call is ONE kernel launch, only enqueued (nothing to report back: the plan checked
        # positive-definiteness on the host); consecutive calls may overlap at their edges (programmatic dependent launch)
        def step(i):
            h.logpdf(mm.desc, ys_dev[i % N_BUF], lml_dev)

        def finish():
            pass
    else:
        sh = sharded.ShardedLogpdf(h, mm, rank, world, dev, overlap=(args.shard_layout == "overlap"))

        def step(i):
            sh.logpdf(ys_dev[i % N_BUF], None, sync=False)      # one launch per shard; the partial lmls land in every rank's buffer

        def finish():
            sh.result(lml_dev)                                  # fixed-order sum over ranks of the LAST step
    c0 = h.counters()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_per_step = timed(step, finish)
    c1 = h.counters()
    value = Tglob / (ms_per_step * 1e-3)
    launches = (c1["launches"] - c0["launches"]) * K // (K + W)
    route = sh.route if sh else "single launch (k_fir_logpdf)"
    # parity of the benchmarked computation at THIS world size: buffer 0 of every rank, concatenated, against the sequential oracle
    step(0)
    finish()
    h.synchronize()
    lml_weak = float(lml_dev.item())

    # ---- the other shard layout, for comparison (N > 1): same series, same kernels, halo received over NVLink vs read in place -----
    alt = None
    if sh is not None and sh.route == "fir":
        alt_layout = "exchange" if sh.overlap else "overlap"
        sh_alt = sharded.ShardedLogpdf(h, mm, rank, world, dev, overlap=(alt_layout == "overlap"))

        def step_alt(i):
            sh_alt.logpdf(ys_dev[i % N_BUF], None, sync=False)

        def finish_alt():
            sh_alt.result(lml_dev)
        ms_alt = timed(step_alt, finish_alt)
        step_alt(0)
        finish_alt()
        h.synchronize()
        alt = {"shard_layout": alt_layout, "ms_per_step": ms_alt, "value": Tglob / (ms_alt * 1e-3), "unit": "steps/s",
               "lml": float(lml_dev.item())}
        assert abs(alt["lml"] - lml_weak) <= 1e-9 * abs(lml_weak), (alt["lml"], lml_weak)

    # ---- e2e: public API, pinned host y, H2D + D2H inside the timed region -------------------------
    with_halo = world > 1 and rank > 0 and sh is not None and sh.overlap
    pin = [torch.from_numpy(np.concatenate([halos[i], v]) if with_halo else v).pin_memory() for i, v in enumerate(ys_host[:2])]
    pin_np = [p.numpy() for p in pin]
    if world == 1:
        def e2e_step(i):
            return pkg.gp.logpdf(fx, pin_np[i % 2])
    else:
        def e2e_step(i):
            return sh.logpdf_host(pin_np[i % 2])
    for i in range(2):
        e2e_step(i)
    barrier()
    ce0 = h.counters()
    t0 = time.perf_counter()
    for i in range(K):
        lml_e2e = e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    ce1 = h.counters()
    e2e_value = Tglob * K / e2e_s
    e2e_h2d = (ce1["h2d_bytes"] - ce0["h2d_bytes"]) / K + (sh.h2d_bytes_per_host_call if sh else 0)
    e2e_d2h = (ce1["d2h_bytes"] - ce0["d2h_bytes"]) / K + (8 if sh else 0)      # sharded: the .item() of the total
    clk = clocks.stop() if rank == 0 else None

    # ---- roofline leg: per-kernel device time with the library's event timers (separate pass, calls do not overlap) ------
    h.set_timing(True)
    for i in range(K):
        step(i)
    tim = h.timing()
    h.set_timing(False)
    tot = sum(t[1] for t in tim) or 1.0
    top = tim[0]
    hbm, peak_src = peaks()
    bytes_per_step = 8.0  # logpdf: y read once, 8 B per Kalman step
    top_alone_ms = top[1] / top[2]
    # The step IS one launch of the dominant kernel, so its average launch duration over the timed region is the timed region's
    # elapsed time / K (consecutive launches overlap at their edges: programmatic dependent launch). The same kernel timed ALONE
    # (events around one launch, the GPU idle before and after: launch latency and the pipeline fill / drain included) is given too.
    one_launch_step = len(tim) == 1 and top[2] == K
    top_ms = ms_per_step if one_launch_step else top_alone_ms
    achieved = bytes_per_step * T / (top_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tp):
        rec = json.load(open(tp)).get(top[0].split("(")[0])
        if rec and rec.get("T") == T:
            traffic = rec["dram_bytes_per_launch"]
    roof = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
            "traffic": traffic, "peak_source": peak_src, "kernel_ms": top_ms, "kernel_ms_timed_alone": top_alone_ms,
            "frac_timed_alone": bytes_per_step * T / (top_alone_ms * 1e-3) / 1e9 / hbm,
            "kernel_share_of_step": top[1] / tot, "algorithmic_bytes_per_step": bytes_per_step,
            "note": "kernel_ms: average duration of the kernel's launches over the timed region (= ms_per_step: one launch per step, "
                    "back to back); kernel_ms_timed_alone: CUDA events around a single launch with the GPU idle on both sides",
            "kernels": [{"name": n, "ms_per_launch": ms / c, "launches_per_step": c / K} for n, ms, c in tim]}

    # ================= strong series (BASELINE config 4): T_total = 8e7 FIXED, sharded over the GPUs ==============================
    strong = None
    if not args.no_strong and STRONG_T % (world * STRONG_CHUNK) == 0:
        del ys_dev
        torch.cuda.empty_cache()
        per = STRONG_T // world
        cpr = per // STRONG_CHUNK                                   # chunks per rank
        mm4 = pkg.lgssm._Marshalled(f(pkg.RegularSpacing(0.0, DT, per), SIGMA2).build_lgssm())
        def series4(b):      # [halo | this rank's chunks]: the overlapped scatter of config 4's series
            parts = [torch.from_numpy(strong_chunk(rank * cpr - 1, b)[-HALO:]) if rank > 0 else torch.zeros(HALO, dtype=torch.float64)]
            parts += [torch.from_numpy(strong_chunk(rank * cpr + c, b)) for c in range(cpr)]
            return torch.cat(parts).to(dev)
        ys4_pad = [series4(b) for b in range(2)]
        ys4 = [p[HALO:] for p in ys4_pad]
        if world == 1:
            def step4(i):
                h.logpdf(mm4.desc, ys4[i % 2], lml_dev)

            def finish4():
                pass
        else:
            sh4 = sharded.ShardedLogpdf(h, mm4, rank, world, dev, overlap=(args.shard_layout == "overlap"))

            def step4(i):
                sh4.logpdf(ys4[i % 2], None, sync=False)

            def finish4():
                sh4.result(lml_dev)
        ms4 = timed(step4, finish4)
        step4(0)
        finish4()
        h.synchronize()
        strong = {"T_total": STRONG_T, "T_per_gpu": per, "value": STRONG_T / (ms4 * 1e-3), "unit": "steps/s", "ms_per_step": ms4,
                  "scaling": "strong", "hbm_frac_per_gpu": 8.0 * per / (ms4 * 1e-3) / 1e9 / hbm, "lml": float(lml_dev.item()),
                  "note": "BASELINE config 4: the SAME series of 8e7 steps at every world size (2 rotating buffer sets); "
                          "scaling = value(N) / value(1) across the lines of a --gpus sweep"}
        del ys4, ys4_pad
        torch.cuda.empty_cache()
        ys_dev = [torch.from_numpy(v).to(dev) for v in ys_host]

    # ---- extra: tgp_filter emitting (m_f, P_f) — 104 B/step algorithmic (single GPU only) --------------
    extra = None
    if world == 1 and not args.no_filter:
        mf = torch.empty((T, 3), dtype=torch.float64, device=dev)
        Pf = torch.empty((T, 9), dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            h.filter(mm.desc, ys_dev[i % N_BUF], mf, 3, Pf, 9, lml_dev)
        torch.cuda.synchronize()
        e0.record(stream)
        for i in range(K):
            h.filter(mm.desc, ys_dev[i % N_BUF], mf, 3, Pf, 9, lml_dev)
        e1.record(stream)
        torch.cuda.synchronize()
        fms = e0.elapsed_time(e1) / K
        fa = 104.0 * T / (fms * 1e-3) / 1e9
        extra = {"value": T / (fms * 1e-3), "unit": "steps/s", "ms_per_step": fms,
                 "roofline": {"bound": "hbm", "achieved": fa, "peak": hbm, "unit": "GB/s", "frac": fa / hbm,
                              "algorithmic_bytes_per_step": 104.0, "note": "whole call (all kernels), y read + (m_f,P_f) written"}}
        del mf, Pf

    secondary = None
    if world == 1 and not args.no_secondary:
        try:
            secondary = secondary_configs(pkg, h, torch)
        except Exception as exc:      # noqa: BLE001 — extras must never take the headline down
            secondary = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- cpu baseline + parity (rank 0, bounded sample) -----------------------------------------------------
    cpu = None
    parity = {}
    if not args.no_cpu:
        from oracle import c_oracle, tgp_oracle as O
        reps = 5
        t_cpu, _ = cpu_reference(T, reps)
        ncores = os.cpu_count() or 1
        all_v, all_n = cpu_all_cores(min(T, 4_000_000), ncores)
        cpu = {"value": T / t_cpu, "unit": "steps/s", "cores": 1, "kind": "port",
               "sample": f"{reps} x full T={T} logpdf with the C restatement of the reference's sequential SArrayStorage path "
                         f"(oracle/lgssm_ref.c, gcc -O3), 1 thread of {ncores} host cores",
               "all_cores": {"value": all_v, "unit": "steps/s", "cores": all_n,
                             "sample": f"{all_n} independent logpdf evaluations of T={min(T, 4_000_000)} each, one POSIX thread per "
                                       "evaluation ('replicas': the only way the single-threaded reference uses every core; NOT "
                                       "what it does for one series)"}}
        # parity at THIS world size, on the benchmarked inputs (buffer 0 of every rank = one series of world * T steps)
        y_all = np.concatenate([synth_y(T, seed_of(r, 0)) for r in range(world)])
        ref = c_oracle.logpdf(c_oracle.Model.from_lgssm(O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, DT, Tglob), SIGMA2)), y_all)
        parity["lml_rel_err_vs_oracle"] = abs(lml_weak - ref) / abs(ref)
        assert parity["lml_rel_err_vs_oracle"] < 1e-6, f"GPU logpdf {lml_weak} differs from the oracle {ref} at world {world}"
        cpu["lml_rel_err_vs_gpu"] = parity["lml_rel_err_vs_oracle"]
        if strong is not None:
            y4 = np.concatenate([strong_chunk(c, 0) for c in range(STRONG_T // STRONG_CHUNK)])
            ref4 = c_oracle.logpdf(c_oracle.Model.from_lgssm(O.build_lgssm(O.Matern52(), O.RegularSpacing(0.0, DT, STRONG_T), SIGMA2)), y4)
            strong["lml_rel_err_vs_oracle"] = abs(strong["lml"] - ref4) / abs(ref4)
            assert strong["lml_rel_err_vs_oracle"] < 1e-6, f"config 4: GPU logpdf {strong['lml']} differs from the oracle {ref4}"

    cfg = workload_config(T, world, args.algo)
    cfg["route"] = route
    if sh is not None:
        cfg["shard_layout"] = ("overlap: every shard is stored with the 3072 observations before it, no inter-GPU dependency inside a step"
                               if sh.overlap else "exchange: the 3072-observation halo is pushed over NVLink by the previous rank's kernel")
    out = {
        "metric": METRIC, "value": value, "unit": "steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": value / 5e7, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "logpdf_per_s": 1e3 / ms_per_step,
        "vs_baseline_note": "BASELINE.md: reference static-lgssm logpdf at N=1e7 read off a plot as 2.5-5e7 steps/s on an unstated CPU; "
                            "5e7 (upper end) used as the denominator",
        "e2e": {"value": e2e_value, "unit": "steps/s", "ms_per_step": e2e_s / K * 1e3,
                "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h,
                "api": "gp.logpdf(to_sde(GP(Matern52Kernel()))(RegularSpacing, sigma2), y_pinned_host)" if world == 1 else
                       "ShardedLogpdf.logpdf_host(y_pinned_host): torch H2D copy of the shard + tgp_shard_logpdf + tgp_shard_result + .item()"},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
        "lml_rel_err_vs_oracle": parity.get("lml_rel_err_vs_oracle"),
        "strong_scaling": strong,
        "other_shard_layout": alt,
        "filter_emit": extra,
        "secondary": secondary,
        "clocks": clk,
        "lml": lml_weak,
    }
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def secondary_configs(pkg, h, torch):
    """BASELINE configs 3 and 5 at full size on this GPU (device-resident inputs, synchronous calls, wall clock around
    torch.cuda.synchronize): reported beside the headline, not part of it. Parity of these paths is the GPU test-suite's job
    (tests/test_gpu_steady_smoother.py, tests/test_gpu_tensorcore.py); details and CPU-oracle times: tools/bench_configs.py,
    tools/cfg5_bench.py, profiles/."""
    out = {}

    def timeit(fn, n, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    hbm, _ = peaks()
    # config 3: D = 10 sum kernel, T = 1e6, posterior marginals at the training inputs (filter + RTS smoother)
    T3 = 1_000_000
    TK = pkg.gp.TransformedKernel
    k3 = 1.0 * pkg.Matern32Kernel() + 0.7 * pkg.Matern52Kernel() + 0.5 * TK(pkg.Matern52Kernel(), 0.5) + 0.3 * TK(pkg.Matern32Kernel(), 2.0)
    m3 = pkg.lgssm._Marshalled(pkg.to_sde(pkg.GP(k3))(pkg.RegularSpacing(0.0, 0.01, T3), 0.1).build_lgssm())
    rng = np.random.default_rng(20261017 + 3)
    y3 = torch.from_numpy(np.sin(np.arange(T3) * 0.004) + 0.3 * np.cos(np.arange(T3) * 0.05) + 0.35 * rng.standard_normal(T3)).cuda()
    Rn = torch.full((1,), 1e-2, dtype=torch.float64, device="cuda")
    md = torch.empty(T3, dtype=torch.float64, device="cuda")
    vd = torch.empty(T3, dtype=torch.float64, device="cuda")
    t3 = timeit(lambda: h.posterior_marginals(m3.desc, y3, Rn, 0, md, vd, None), 10, 3)
    out["cfg3_posterior_marginals_D10_T1e6"] = {"ms": t3 * 1e3, "steps_per_s": T3 / t3, "dtype": "f64",
                                                "hbm_frac_of_measured": 1784.0 * T3 / t3 / 1e9 / hbm, "algorithmic_bytes_per_step": 1784.0}
    del y3, md, vd
    # config 5: Separable(SE, Matern52), 256 spatial points x T = 1e5 (D = 768, M = 256), FP32 storage on the tensor cores, logpdf
    Nr, T5 = 256, 100_000
    r = np.linspace(-3.0, 3.0, Nr)
    fx5 = pkg.to_sde(pkg.GP(pkg.Separable(pkg.SEKernel(), pkg.Matern52Kernel())), pkg.ArrayStorage(np.float32))(
        pkg.RectilinearGrid(r, pkg.RegularSpacing(0.0, 0.01, T5)), 0.1)
    m5 = pkg.lgssm._Marshalled(fx5.build_lgssm())
    y5 = torch.from_numpy(np.random.default_rng(20261017 + 5).standard_normal((T5, Nr))).cuda()
    lml5 = np.zeros(1)
    h5 = pkg.Handle(h.device)    # its own handle and stream, switched to the FP32-storage tensor-core arithmetic
    h5.set_dense_math(pkg.lgssm.TGP_DENSE_TF32X3)
    try:
        t5 = timeit(lambda: h5.logpdf(m5.desc, y5, lml5), 2, 1)
    finally:
        h5.set_dense_math(pkg.lgssm.TGP_DENSE_F64)
    out["cfg5_logpdf_D768_M256_T1e5_fp32_tensorcore"] = {"s": t5, "steps_per_s": T5 / t5, "dtype": "f32 storage, 3xTF32 tcgen05, FP64 Cholesky / means",
                                                         "lml": float(lml5[0])}
    # the algorithm north_star names, un-specialised: config 2 forced through the general 5-tuple scan, and a time-varying model
    try:
        from tools import bench_general
        out.update(bench_general.run(pkg, h, torch))
    except Exception as exc:      # noqa: BLE001
        out["general_scan_error"] = f"{type(exc).__name__}: {exc}"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--T", type=int, default=T_DEFAULT)
    ap.add_argument("--algo", default="auto", choices=["auto", "scan"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-filter", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs 3 / 5 extras")
    ap.add_argument("--shard-layout", choices=["overlap", "exchange"], default="overlap",
                    help="N > 1: how the series is scattered (overlap = each shard carries its 3072-observation halo)")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-T = 8e7 (config 4) series")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
