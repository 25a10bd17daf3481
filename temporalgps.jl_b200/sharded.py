"""
sharded.py — time-sharded logpdf over several GPUs, one process per GPU.

Rank r owns steps [b[r], b[r+1]) of ONE series. Three routes, chosen COLLECTIVELY (every rank runs the same one):

  "fir"      time-invariant models within the range of tgp_shard_logpdf (the path BASELINE config 4 runs): one kernel launch per
             shard and call, nothing else. Rank 0 runs the covariance transient; rank r > 0 starts from the <= 3072 observations
             that precede its shard, which rank r - 1's kernel stores into r's exchange buffer over NVLink when it STARTS, so no
             shard ever waits for another shard's result. The partial log-likelihoods land in every rank's buffer; `result()`
             sums them (fixed order) when the caller wants the number.
  "steady"   time-invariant models the first route declines (filters with a long memory): tgp_shard_phase1 -> all_gather of one
             affine record (D*D + D doubles) per rank -> tgp_shard_phase2 -> all_reduce, on ONE stream, no host round trip.
  "general"  everything else: tgp_shard_reduce folds the shard into one scan element (3D^2 + 2D doubles), all_gather, every rank
             folds the elements before it into x0 (tgp_shard_prefix), ordinary tgp_logpdf from that state, all_reduce.

Route "fir" has a second data layout, `overlap=True`: the series is scattered so that every shard with rank > 0 is preceded IN ITS
OWN device buffer by the TGP_SHARD_HALO (3072) observations before it (shard_with_halo below). The kernel then reads them in place —
nothing is pushed between GPUs, no rank ever waits for another, and the only traffic left is the 16-byte partial result per peer.

posterior_marginals_sharded: the smoother sharded the same way (halos on both sides, no exchange at all).

The reference has no analogue (single-threaded); host logic only here — arithmetic is in the library.
"""
from __future__ import annotations

import numpy as np

from ._lib import TGP_SHARD_HALO  # noqa: F401  (re-exported: the overlapped layout's halo length)


def shard_bounds(T_total: int, world: int):
    """Contiguous, near-equal time shards: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(int(T_total), int(world))
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def shard_with_halo(torch, y_full_host, bounds, rank, device):
    """The overlapped scatter: -> (buffer, shard view). buffer = [halo | shard] on `device`; the view is what logpdf() takes."""
    from ._lib import TGP_SHARD_HALO
    lo, hi = bounds[rank], bounds[rank + 1]
    halo = TGP_SHARD_HALO if rank > 0 else 0
    if lo - halo < 0:
        raise ValueError(f"shard {rank} starts at {lo}: fewer than {halo} observations precede it")
    buf = torch.empty(TGP_SHARD_HALO + hi - lo, dtype=torch.float64, device=device)   # same offset on every rank: 32-byte aligned view
    buf[TGP_SHARD_HALO - halo:].copy_(torch.from_numpy(np.ascontiguousarray(y_full_host[lo - halo:hi])))
    return buf, buf[TGP_SHARD_HALO:]


def halo_bounds(T_total, world, rank, halo):
    """-> (lo, hi, lo_h, hi_h): the shard [lo, hi) of `rank` and the same extended by `halo` steps on both sides (clipped)."""
    b = shard_bounds(T_total, world)
    lo, hi = b[rank], b[rank + 1]
    return lo, hi, max(0, lo - halo), min(int(T_total), hi + halo)


def posterior_marginals_sharded(gp, f, x, noise, y_ext, noise_new, rank, world, halo=4096):
    """Time-sharded marginals(posterior(f(x, noise), y)(x, noise_new)) at the training inputs of ONE regular-grid series
    (posterior_lti_sde.jl:27-36), with NO exchange between the ranks: rank r filters and smooths its shard extended by `halo` steps on
    both sides and keeps the interior. A stable filter forgets its start and the RTS recursion forgets its end at the same rate, so
    the interior equals the global posterior to |Abar^halo| (config 3: 1e-30 at halo = 4096); the extended shard starts from the
    stationary prior and ends without a successor exactly like the first and last shards of the series do.
    gp: the mirror module (temporalgps.jl_b200.gp); f: LTISDE; x: RegularSpacing of the WHOLE series; y_ext: this rank's observations
    [lo_h, hi_h) (halo_bounds); noise / noise_new: scalars. -> (mean, var) of the rank's own steps [lo, hi)."""
    lo, hi, lo_h, hi_h = halo_bounds(len(x), world, rank, halo)
    if len(y_ext) != hi_h - lo_h:
        raise ValueError(f"rank {rank}: y_ext must hold the {hi_h - lo_h} observations [{lo_h}, {hi_h}), got {len(y_ext)}")
    x_ext = type(x)(x.t0 + lo_h * x.dt, x.dt, hi_h - lo_h)
    mu, var = gp.marginals(gp.posterior(f(x_ext, noise), y_ext)(x_ext, noise_new))
    return mu[lo - lo_h:hi - lo_h], var[lo - lo_h:hi - lo_h]


def incoming_state(prefix_fn, D, elems, rank, m0, P0):
    """State entering rank's shard from the gathered elements (rows 0..rank-1)."""
    return prefix_fn(D, elems[:rank] if rank else None, m0, P0)


def agree(dist, world, flag: bool) -> bool:
    """True iff `flag` is true on every rank (the route must be the same everywhere, or the collectives mismatch)."""
    if world == 1 or dist is None:
        return bool(flag)
    flags = [None] * world
    dist.all_gather_object(flags, bool(flag))
    return all(flags)


class ShardedLogpdf:
    def __init__(self, handle, marshalled, rank, world, device, dist=None, route=None, overlap=False):
        import torch
        self.torch = torch
        if dist is None:
            import torch.distributed as dist
        self.dist = dist
        self.h, self.mm, self.rank, self.world, self.dev = handle, marshalled, rank, world, device
        # The library's kernels and torch's collectives must be ordered on ONE stream. torch's default stream has handle 0, which
        # tgp_set_stream reads as "use the handle's own stream" — so run everything on a real side stream and make it current.
        if getattr(device, "type", "cpu") == "cuda":
            st = torch.cuda.current_stream(device)
            if st.cuda_stream == 0:
                st = torch.cuda.Stream(device=device)
                torch.cuda.set_stream(st)
            self.stream = st
            handle.set_stream(st.cuda_stream)
        self.D = marshalled.D
        self.ES = 3 * self.D * self.D + 2 * self.D
        self.XS = self.D * self.D + self.D
        self.elem = torch.zeros(self.ES, dtype=torch.float64, device=device)
        self.all = torch.zeros(world * self.ES, dtype=torch.float64, device=device)
        self.part = torch.zeros(1, dtype=torch.float64, device=device)
        self.rec = torch.zeros(self.XS, dtype=torch.float64, device=device)
        self.recs = torch.zeros(world * self.XS, dtype=torch.float64, device=device)
        self.m0 = np.array(marshalled.keep[-2])
        self.P0 = np.array(marshalled.keep[-1])
        self.desc2 = type(marshalled.desc).from_buffer_copy(marshalled.desc)   # ctypes structs with pointers cannot be copy.copy'd
        self._ybuf = None
        self._deferred_on = False
        d = marshalled.desc
        ti = not (d.sA or d.sa or d.sQ or d.sH or d.sh or d.sR) and d.ordering == 0
        # ---- route, decided collectively --------------------------------------------------------------------------------
        self.transport_error = None
        self.route = "general"
        cuda = getattr(device, "type", "cpu") == "cuda"
        if route in (None, "fir") and cuda and agree(dist, world, ti and self.D <= 4 and d.M == 1 and d.T >= 4096):
            ok = True
            if world > 1:
                ok = self._open_exchange()
            if ok:
                self.route = "fir"
        if self.route != "fir" and route in (None, "fir", "steady") and agree(dist, world, ti and d.T >= 65536 and self.D <= 6):
            self.route = "steady"
        if route == "general":
            self.route = "general"
        self.overlap = bool(overlap) and self.route == "fir" and world > 1

    def _open_exchange(self) -> bool:
        """CUDA IPC handles of the ranks' exchange buffers, gathered once. False (on every rank) if peer memory is unavailable."""
        h, dist, world = self.h, self.dist, self.world
        if getattr(h, "_xchg_opened", None) == (self.rank, world):   # one exchange per handle, shared by every ShardedLogpdf on it
            return agree(dist, world, True)
        mine, err = None, None
        try:
            mine = h.xchg_create(self.rank, world, 16)
        except Exception as exc:      # noqa: BLE001 — reported in transport_error, another route is used instead
            err = str(exc)
        handles = [None] * world
        dist.all_gather_object(handles, mine)          # once per run: 64-byte CUDA IPC handles, any backend
        ok = False
        if all(hd is not None for hd in handles):
            try:
                h.xchg_open(b"".join(handles))
                ok = True
            except Exception as exc:  # noqa: BLE001
                err = str(exc)
        self.transport_error = err
        ok = agree(dist, world, ok)
        if ok:
            h._xchg_opened = (self.rank, world)
        return ok

    # ------------------------------------------------------------------------------------------------------------------
    def logpdf(self, y_dev, lml_out_dev=None, sync=True):
        """y_dev: this rank's shard, resident on its GPU. lml_out_dev: 1-element CUDA tensor receiving the total (all ranks), or
        None to leave the total to a later result() call (route "fir": the step is then exactly one kernel launch).
        sync=False: nothing waits on the host; failures surface at check()."""
        h, dist = self.h, self.dist
        if self.route == "fir":
            if getattr(h, "_shard_overlap", None) != self.overlap:     # several ShardedLogpdf objects may share the handle
                from ._lib import TGP_OPT_SHARD_OVERLAP
                h.set_option(TGP_OPT_SHARD_OVERLAP, 1 if self.overlap else 0)
                h._shard_overlap = self.overlap
            try:
                h.shard_logpdf(self.mm.desc, y_dev, self.rank, self.world)
            except Exception as exc:                     # the plan declined this model (the same decision on every rank)
                if getattr(exc, "code", None) != 5:      # TGP_EUNSUPPORTED
                    raise
                d = self.mm.desc
                self.route = "steady" if (d.T >= 65536 and self.D <= 6) else "general"
                return self.logpdf(y_dev, lml_out_dev, sync)
            if lml_out_dev is not None:
                h.shard_result(lml_out_dev)
            if sync:
                h.synchronize()
            return
        if lml_out_dev is None:
            lml_out_dev = self.part
        if self.route == "steady":
            if not self._deferred_on:
                from ._lib import TGP_OPT_DEFER_STATUS
                h.set_option(TGP_OPT_DEFER_STATUS, 1)
                self._deferred_on = True
            h.shard_phase1(self.mm.desc, y_dev, self.rank, self.world, self.rec)
            dist.all_gather_into_tensor(self.recs, self.rec)
            h.shard_phase2(self.recs, lml_out_dev)     # enqueued, not synchronised
            dist.all_reduce(lml_out_dev)
            if sync:
                h.synchronize()                         # status of the shard (convergence, positive-definiteness)
            return
        h.shard_reduce(self.mm.desc, y_dev, self.elem)
        dist.all_gather_into_tensor(self.all, self.elem)
        elems = self.all.cpu().numpy().reshape(self.world, self.ES)
        m_in, P_in = incoming_state(h.shard_prefix, self.D, elems, self.rank, self.m0, self.P0)
        self._keep = (np.ascontiguousarray(m_in), np.ascontiguousarray(P_in))
        self.desc2.m0 = self._keep[0].ctypes.data
        self.desc2.P0 = self._keep[1].ctypes.data
        part = self.torch.zeros(1, dtype=self.torch.float64, device=self.dev)
        h.logpdf(self.desc2, y_dev, part)
        dist.all_reduce(part)
        lml_out_dev.copy_(part)

    def result(self, lml_out_dev):
        """Total log-likelihood of the LAST logpdf() call into a 1-element CUDA tensor (stream-ordered)."""
        if self.route == "fir":
            self.h.shard_result(lml_out_dev)
        else:
            lml_out_dev.copy_(self.part)

    def check(self):
        """Wait for the stream and raise if any un-synchronised call since the last check failed."""
        self.h.synchronize()

    def close(self):
        if self._deferred_on:
            from ._lib import TGP_OPT_DEFER_STATUS
            self.h.set_option(TGP_OPT_DEFER_STATUS, 0)
            self._deferred_on = False

    def logpdf_host(self, y_host_pinned):
        """End-to-end variant: the shard's observations start in pinned host memory. With the overlapped layout the host array of a
        rank > 0 is [the TGP_SHARD_HALO observations before the shard | the shard]."""
        from ._lib import TGP_SHARD_HALO
        torch = self.torch
        n = len(y_host_pinned)
        if self._ybuf is None or self._ybuf.numel() != n:
            self._ypad = torch.empty(TGP_SHARD_HALO + n, dtype=torch.float64, device=self.dev)
            self._ybuf = self._ypad[TGP_SHARD_HALO:]
            self._out = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self._ybuf.copy_(torch.from_numpy(y_host_pinned), non_blocking=True)
        halo = TGP_SHARD_HALO if (self.overlap and self.rank > 0) else 0
        self.logpdf(self._ybuf[halo:], self._out)
        return float(self._out.item())

    @property
    def h2d_bytes_per_host_call(self):
        """Bytes logpdf_host moves to the device per call through torch (not counted by the library's own counters)."""
        return 0 if self._ybuf is None else self._ybuf.numel() * 8
