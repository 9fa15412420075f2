"""Multi-GPU decision + PCS: one process per GPU, ``torch.distributed`` for the (tiny) exchanges.

Two partitions of the arm x context grid are supported:

``contexts`` (default)
    every rank holds the whole (small, ~MBs) GP state and owns a contiguous slice of the context
    set.  Grid posterior, scan, step 2(b) and PCS counts are purely local; the only exchange of a
    decision is ONE all-gather of the 64-byte ``crs_decision`` of each rank's local minimiser, and
    the PCS estimate needs one all-gather of the per-context counts (or a scalar all-reduce).
``arms`` (the partition BASELINE.json's north_star describes)
    rank g owns arms [g*K/G, (g+1)*K/G) for ALL contexts.  After the local grid, ranks all-gather
    their per-context local best (mean, variance, arm id), reduce it to the global best per
    context, scan their own columns against it (``ext_best_*`` of ``crs_ocba_scan``), all-gather
    the local minimisers, and all-reduce the two psi partial sums of step 2(b).

The combination rules are pure functions of the gathered buffers so that they can be tested on
the CPU with the gloo backend (tests/test_sharded_cpu.py).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib

# column layout of the float64 record every rank contributes to the decision all-gather
REC_MIN_ZETA, REC_TIES, REC_CONTEXT, REC_ZETA_ARM, REC_BEST_ARM, REC_NEXT_ARM, REC_PSI_BEST, REC_PSI_REST = range(8)
REC_WIDTH = 8


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of ``total`` items for ``rank`` of ``world``."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def decision_record(min_zeta: float, ties: int, context: int, zeta_arm: int, best_arm: int,
                    next_arm: int, psi_best: float = 0.0, psi_rest: float = 0.0) -> Tensor:
    """One rank's local minimiser as a float64 row (indices < 2^53 are exact in float64)."""
    return torch.tensor([min_zeta, ties, context, zeta_arm, best_arm, next_arm, psi_best, psi_rest],
                        dtype=torch.float64)


def records_from_bytes(raw: Tensor, n_c: int, world: int) -> Tensor:
    """``[world, 64]`` uint8 crs_decision structs (LOCAL context indices of a context-sharded
    run) -> ``[world, REC_WIDTH]`` float64 records with GLOBAL context indices."""
    out = torch.zeros(world, REC_WIDTH, dtype=torch.float64)
    buf = raw.numpy()
    for r in range(world):
        d = _lib.Decision.from_buffer_copy(buf[r].tobytes())
        begin, _ = shard_bounds(n_c, world, r)
        out[r] = decision_record(d.min_zeta, d.tie_count, d.context + begin if d.context >= 0 else -1,
                                 d.zeta_arm, d.best_arm, d.next_arm, d.psi_best, d.psi_rest)
    return out


def combine_context_sharded(records: Tensor) -> Tensor:
    """``records``: ``[world, REC_WIDTH]`` (contexts already GLOBAL indices).  The global decision
    is the record with the smallest zeta; equal zeta -> lowest global context (the reference's
    first-minimum rule in flat context-major order); the tie count is summed over ranks that attain
    the minimum.  Ranks with no contexts contribute min_zeta = +inf."""
    z = records[:, REC_MIN_ZETA]
    zmin = z.min()
    at_min = z == zmin
    ctx = torch.where(at_min, records[:, REC_CONTEXT], torch.full_like(z, float("inf")))
    winner = int(torch.argmin(ctx))
    out = records[winner].clone()
    out[REC_TIES] = records[at_min, REC_TIES].sum()
    return out


def combine_local_best(means: Tensor, variances: Tensor, arms: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Arm sharding, step 2: ``[world, n_c]`` gathered per-context local bests -> global best per
    context (largest mean, lowest arm id on ties)."""
    top = means.max(dim=0).values
    cand = torch.where(means == top.unsqueeze(0), arms, torch.full_like(arms, torch.iinfo(arms.dtype).max))
    arm = cand.min(dim=0).values
    owner = (cand == arm.unsqueeze(0)).to(torch.int64).argmax(dim=0)
    idx = owner.unsqueeze(0)
    return means.gather(0, idx).squeeze(0), variances.gather(0, idx).squeeze(0), arm


def combine_arm_sharded(records: Tensor, mean_at: Optional[Tensor] = None) -> Tensor:
    """Arm sharding, step 3/4 (host restatement of ``crs_arm_fold``): every rank's record refers to the
    SAME context set; the global minimiser is the smallest zeta, then the lowest context, then — the
    same context on several shards — the zeta arm with the LARGER mean (``mean_at[r]``, the reference's
    descending-mean sort order, which is also the single-GPU kernel's rule), then the lower arm id.
    Tie counts of all records at the minimum are added."""
    z = records[:, REC_MIN_ZETA]
    zmin = z.min()
    at_min = z == zmin
    ctx = torch.where(at_min, records[:, REC_CONTEXT], torch.full_like(z, float("inf")))
    cmin = ctx.min()
    both = at_min & (records[:, REC_CONTEXT] == cmin)
    if mean_at is not None:
        mtop = torch.where(both, mean_at.to(z.dtype), torch.full_like(z, -float("inf"))).max()
        both = both & (mean_at.to(z.dtype) == mtop)
    arms = torch.where(both, records[:, REC_ZETA_ARM], torch.full_like(z, float("inf")))
    winner = int(torch.argmin(arms))
    out = records[winner].clone()
    out[REC_TIES] = records[at_min, REC_TIES].sum()
    return out


def all_gather_rows(row: Tensor, group=None) -> Tensor:
    """All-gather one equal-shaped tensor per rank into ``[world, ...]`` (NCCL on GPUs, gloo on CPU)."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(row.shape), dtype=row.dtype, device=row.device)
    dist.all_gather_into_tensor(out, row.contiguous(), group=group) if row.is_cuda else \
        dist.all_gather(list(out.unbind(0)), row.contiguous(), group=group)
    return out


class PeerMailbox:
    """Peer-mapped mailboxes of the ranks of ``group`` (one process per GPU of one NVLink domain):
    the device-side exchange ``crs_ocba_scan_select_sharded`` / ``crs_peer_sum_u64`` use instead of an
    NCCL launch + host parse (include/crs_b200.h ``crs_peer_exchange``).  Set-up (once): every rank
    allocates its mailbox, the 64-byte CUDA IPC handles are all-gathered through ``group`` and every
    peer's mailbox is opened.  Every rank must issue the same sequence of exchanges."""

    def __init__(self, device, group=None):
        self.lib = _lib.get()
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.seq = 0
        self._own = C.c_void_p()
        self._opened: List[C.c_void_p] = []
        self._ptrs = None
        if self.world == 1:
            return
        # Every rank runs the SAME sequence of collectives whatever fails locally (no rank may be left waiting
        # in a barrier); the outcome is agreed on with one all-reduce, and on failure every rank raises.
        err = None
        with torch.cuda.device(self.device):
            handle = (C.c_ubyte * 64)()
            try:
                import os
                if os.environ.get("CRS_DISABLE_IPC") == "1":  # test hook: this rank cannot map peer memory
                    raise RuntimeError("CUDA IPC disabled by CRS_DISABLE_IPC")
                _lib.check(self.lib.crs_device_alloc(int(self.lib.crs_mailbox_bytes(self.world)), C.byref(self._own)),
                           "crs_device_alloc")
                _lib.check(self.lib.crs_ipc_export(self._own, handle), "crs_ipc_export")
            except Exception as e:  # noqa: BLE001
                err = e
            mine = torch.frombuffer(bytearray(bytes(handle)), dtype=torch.uint8).to(self.device)
            every = torch.empty(self.world, 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(every, mine, group=group)
            every = every.cpu()
            ptrs = []
            if err is None:
                try:
                    for r in range(self.world):
                        if r == self.rank:
                            ptrs.append(self._own.value)
                            continue
                        peer = C.c_void_p()
                        buf = (C.c_ubyte * 64).from_buffer_copy(every[r].numpy().tobytes())
                        _lib.check(self.lib.crs_ipc_open(buf, C.byref(peer)), "crs_ipc_open")
                        self._opened.append(peer)
                        ptrs.append(peer.value)
                except Exception as e:  # noqa: BLE001
                    err = e
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also: nobody posts before everybody has mapped everybody
            if int(ok) == 0:
                self._release()
                raise RuntimeError("peer mailboxes unavailable (CUDA IPC / peer access failed on at least one rank"
                                   + (f"; here: {err}" if err is not None else "") + ")")
            self._ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)

    def _release(self) -> None:
        for p in self._opened:
            self.lib.crs_ipc_close(p)
        self._opened = []
        if self._own.value is not None:
            self.lib.crs_device_free(self._own)
            self._own = C.c_void_p()

    def next_exchange(self, ctx_begin: int = 0) -> "_lib.PeerExchange":
        """The descriptor of the next exchange; the sequence number advances by one on every rank."""
        self.seq = self.seq % 0x3ffffffe + 1  # 1 .. 0x3ffffffe: the slot parity (seq & 1) keeps alternating across the wrap
        return _lib.PeerExchange(self._ptrs.data_ptr() if self._ptrs is not None else None, self.rank, self.world,
                                 self.seq, 0, int(ctx_begin))

    def close(self) -> None:
        if self._own.value is None and not self._opened:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)  # no peer is still writing into a mailbox about to go away
            self._release()

    def __del__(self):  # pragma: no cover
        try:
            for p in self._opened:
                self.lib.crs_ipc_close(p)
            if self._own.value is not None:
                self.lib.crs_device_free(self._own)
        except Exception:
            pass


class ContextShardedDecider:
    """GP-C-OCBA decision (+ PCS) with the context set sharded over the ranks of ``group``.

    Every rank must construct it with the same model description and the same full
    ``context_set``; rank r keeps rows ``shard_bounds(n_c, world, r)``.

    ``exchange="mailbox"`` (default on GPUs): the ranks' local minimisers are combined ON THE DEVICE by
    the kernel that finishes the local decision (``crs_ocba_scan_select_sharded``: 64-byte stores into
    the peers' mailboxes over NVLink, fold in the last CTA, result in mapped pinned host memory) and the
    PCS totals by ``crs_peer_sum_u64`` — no NCCL launch, no stream synchronise, no Python parsing.
    ``exchange="nccl"``: one 64-byte-per-rank ``all_gather`` + host-side fold (round-1 path, kept for
    comparison and for groups without peer access).
    """

    def __init__(self, model, context_set: Tensor, group=None, exchange: str = "mailbox"):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_c = context_set.shape[0]
        self.begin, self.end = shard_bounds(self.n_c, self.world, self.rank)
        self.local_ctx = context_set[self.begin:self.end].to(device=model.device, dtype=model.dtype).contiguous()
        # raw 64-byte crs_decision structs are exchanged as bytes; an empty shard sends +inf
        empty = _lib.Decision(min_zeta=float("inf"), context=-1, zeta_arm=-1, best_arm=-1, next_arm=-1)
        self._empty = torch.frombuffer(bytearray(bytes(empty)), dtype=torch.uint8).to(model.device)
        self._gather = torch.zeros(self.world, 64, dtype=torch.uint8, device=model.device)
        self._gather_host = torch.zeros(self.world, 64, dtype=torch.uint8).pin_memory()
        # the decider OWNS its workspace: the cached local grid must survive any number of stateless
        # calls of any batch size passing through the same model
        self.n_loc = self.end - self.begin
        self.ws = model.private_workspace(max(self.n_loc, 1))
        if exchange not in ("mailbox", "nccl"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self.exchange = exchange if self.world > 1 else "nccl"
        self.mailbox = None
        if self.exchange == "mailbox":
            try:
                self.mailbox = PeerMailbox(model.device, group)
            except RuntimeError as e:  # raised on EVERY rank (the outcome is agreed on collectively): fall back together
                import sys
                print(f"contextual_rs_b200: {e}; falling back to exchange='nccl'", file=sys.stderr)
                self.exchange = "nccl"
        self._sum_hosts = [torch.zeros(10, dtype=torch.int64).pin_memory() for _ in range(2)]
        self._sum_slot = 0
        self._sum_dev = torch.zeros(8, dtype=torch.int64, device=model.device)
        self._cdf_ws = None

    def _launch_local(self, p_mode: int, kernel_scale: float, prepare_arms, grid: bool, px) -> None:
        """prepare -> local grid -> scan / select / peer combine, all on the current stream."""
        m, lib, ws, n_loc = self.model, _lib.get(), self.ws, self.n_loc
        K = m.num_arms
        stream = _lib.stream_ptr(m.device)
        if prepare_arms is not None and prepare_arms[1] > prepare_arms[0]:
            m.prepare(prepare_arms[0], prepare_arms[1], relearn=False)
        st = C.byref(m.state)
        if n_loc > 0 and grid:
            _lib.check(lib.crs_posterior_grid(st, self.local_ctx.data_ptr(), n_loc, 0, K, ws["mean"].data_ptr(),
                                              ws["var"].data_ptr(), K, 0, stream), "crs_posterior_grid")
        _lib.check(lib.crs_ocba_scan_select_sharded(
            st, self.local_ctx.data_ptr(), ws["mean"].data_ptr(), ws["var"].data_ptr(), n_loc, K, p_mode,
            float(kernel_scale), ws["best_arm"].data_ptr(), ws["min_zeta"].data_ptr(), ws["min_arm"].data_ptr(),
            ws["tie_cnt"].data_ptr(), ws["decision"].data_ptr(), ws["scan_ws"].data_ptr(), C.byref(px),
            ws["decision_host"].data_ptr(), stream), "crs_ocba_scan_select_sharded")

    def decide_struct(self, p_mode: int = _lib.CRS_P_COUNT, kernel_scale: float = 2.0,
                      prepare_arms: Optional[Tuple[int, int]] = None, grid: bool = True) -> "_lib.Decision":
        """Mailbox exchange: launch, then poll the mapped host copy of the GLOBAL decision (global
        context index).  ``grid=False`` re-uses the cached local grid (sequential loop).  The returned
        struct is a VIEW of this decider's pinned host buffer: read it before the next decision overwrites it
        (:meth:`decide` returns a copy)."""
        m = self.model
        px = self.mailbox.next_exchange(self.begin)
        with torch.cuda.device(m.device):
            self._launch_local(p_mode, kernel_scale, prepare_arms, grid, px)
            _lib.check(_lib.get().crs_wait_decision(self.ws["decision_host"].data_ptr(), px.seq, 20000), "crs_wait_decision")
        dec = self.ws["decision_view"]
        if dec.reserved[0] == -2:
            raise RuntimeError("peer mailbox exchange timed out: a rank did not post its decision record")
        if dec.reserved[0] > 0:
            raise RuntimeError(f"arm {dec.reserved[0] - 1}: kernel matrix not positive definite")
        return dec

    def decide(self, p_mode: int = _lib.CRS_P_COUNT, kernel_scale: float = 2.0,
               prepare_arms: Optional[Tuple[int, int]] = None, grid: bool = True) -> Tensor:
        """Returns the combined record (``REC_*`` columns, global context index) on the host."""
        from .contextual_rs_strategies import _scan_select
        m, lib = self.model, _lib.get()
        n_loc = self.end - self.begin
        if self.exchange == "mailbox":
            d = self.decide_struct(p_mode, kernel_scale, prepare_arms, grid)
            return decision_record(d.min_zeta, d.tie_count, d.context, d.zeta_arm, d.best_arm, d.next_arm,
                                   d.psi_best, d.psi_rest)
        if n_loc > 0:
            ws = self.ws
            K = m.num_arms
            with torch.cuda.device(m.device):
                stream = _lib.stream_ptr(m.device)
                if prepare_arms is not None and prepare_arms[1] > prepare_arms[0]:
                    m.prepare(prepare_arms[0], prepare_arms[1], relearn=False)
                st = C.byref(m.state)
                if grid:
                    _lib.check(lib.crs_posterior_grid(st, self.local_ctx.data_ptr(), n_loc, 0, K, ws["mean"].data_ptr(),
                                                      ws["var"].data_ptr(), K, 0, stream), "crs_posterior_grid")
                _scan_select(lib, m, ws, n_loc, p_mode, kernel_scale, stream, ctx_ptr=self.local_ctx.data_ptr())
            local = ws["decision"]
        else:
            local = self._empty
        if self.world > 1:
            dist.all_gather_into_tensor(self._gather, local, group=self.group)
            self._gather_host.copy_(self._gather, non_blocking=True)
        else:
            self._gather_host[0].copy_(local, non_blocking=True)
        torch.cuda.current_stream(m.device).synchronize()
        return combine_context_sharded(records_from_bytes(self._gather_host, self.n_c, self.world))

    def pcs_counts(self, num_samples: int, seed: int, offset: int = 0) -> Tensor:
        """Local success counts (global context indices key the stream, so the result does not
        depend on the number of ranks); call ``all_gather_counts`` to assemble the full vector."""
        from .generalized_pcs import pcs_counts
        n_loc = self.end - self.begin
        ws = self.ws
        if self._cdf_ws is None:
            nbytes = int(_lib.get().crs_pcs_cdf_workspace_bytes(max(n_loc, 1)))
            self._cdf_ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.model.device)
        return pcs_counts(self.model, ws["mean"][:n_loc], ws["var"][:n_loc], num_samples, seed, offset,
                          ctx_offset=self.begin, counts=ws["counts"][:n_loc], stats=ws["stats"], sampler="cdf",
                          workspace=self._cdf_ws)

    def pcs_mean(self, num_samples: int, seed: int, offset: int = 0) -> Tensor:
        """rho = mean over contexts: the kernels accumulate the rank's total in ``stats[4]``; the ranks'
        totals are added by ``crs_peer_sum_u64`` (mailbox) or one scalar all-reduce (nccl)."""
        if self.exchange == "mailbox":
            return torch.tensor([self.pcs_total(num_samples, seed, offset) / (float(num_samples) * self.n_c)],
                                dtype=torch.float64)
        total = self.pcs_counts(num_samples, seed, offset).sum(dtype=torch.int64).to(torch.float64).reshape(1)
        if self.world > 1:
            dist.all_reduce(total, group=self.group)
        return total / (float(num_samples) * self.n_c)

    def pcs_total_async(self, num_samples: int, seed: int, offset: int = 0) -> "PendingTotal":
        """Launch the estimate (PCS kernels + the peer sum of the ranks' totals) and return at once; the
        :class:`PendingTotal` is read later (``.result()``), so that host work — the simulation, the update of
        the sampled arm — overlaps the estimate on the GPU.  Two host slots alternate: a result must be read
        before two further estimates are launched."""
        lib, m = _lib.get(), self.model
        stats = self.ws["stats"]
        if self.n_loc:
            self.pcs_counts(num_samples, seed, offset)  # the cdf stream leaves the rank's total in stats[4]
        else:
            stats.zero_()
        px = self.mailbox.next_exchange(self.begin)
        self._sum_slot ^= 1
        host = self._sum_hosts[self._sum_slot]
        with torch.cuda.device(m.device):
            _lib.check(lib.crs_peer_sum_u64(stats.data_ptr() + 4 * 8, 1, self._sum_dev.data_ptr(),
                                            host.data_ptr(), C.byref(px), _lib.stream_ptr(m.device)),
                       "crs_peer_sum_u64")
        return PendingTotal(host, px.seq)

    def pcs_total(self, num_samples: int, seed: int, offset: int = 0) -> int:
        """Sum over ALL contexts of all ranks of the success counts (host int), mailbox exchange."""
        return self.pcs_total_async(num_samples, seed, offset).result()


class PendingTotal:
    """A PCS total on its way to mapped host memory (``crs_peer_sum_u64``): ``result()`` polls the tag."""

    def __init__(self, host: Tensor, seq: int):
        self._host, self._seq, self._value = host, seq, None

    def result(self) -> int:
        if self._value is None:
            _lib.check(_lib.get().crs_wait_u64(self._host.data_ptr() + 8 * 8, self._seq, 20000), "crs_wait_u64")
            if int(self._host[9]) != 0:
                raise RuntimeError("peer mailbox exchange timed out: a rank did not post its PCS total")
            self._value = int(self._host[0])
        return self._value


def sub_description(desc: dict, a0: int, a1: int) -> dict:
    """The fitted model description restricted to arms [a0, a1)."""
    out = dict(desc)
    for key in ("train_X", "train_Y"):
        out[key] = list(desc[key][a0:a1])
    for key in ("lengthscale", "outputscale", "noise", "mean_const", "norm_mins", "norm_ranges", "y_mean", "y_std"):
        if desc.get(key) is not None:
            out[key] = desc[key][a0:a1]
    return out


class ArmShardedDecider:
    """GP-C-OCBA decision (+ PCS) with the ARMS sharded over the ranks — the partition
    BASELINE.json's north_star describes (each arm is an independent GP).

    Rank g owns arms ``shard_bounds(K, world, g)`` for ALL contexts.  One decision is stream-resident —
    no host parsing between its steps; the host reads the finished 64-byte ``crs_decision`` once:

      1. local grid ``[n_c, K/G]`` + local per-context best (``crs_ocba_scan``) -> ``crs_arm_pack_best``
      2. **all-gather** of the packed bests, ``3 x n_c`` doubles per rank (NCCL; C4: 1.5 MB/rank)
         -> ``crs_arm_combine_best`` = global best per context
      3. local zeta minimiser against the global best (``crs_ocba_scan`` with ``ext_best_*``)
         -> ``crs_arm_record_mean`` -> **all-gather** of the 64-byte records -> ``crs_arm_fold``
      4. psi partial sums of step 2(b) at the selected context (``crs_ocba_select`` on the local arms)
         -> **all-reduce** of 2 doubles (+ status) -> ``crs_arm_finish``.

    PCS needs every arm of a context, so the grid is re-sharded by context with one all-to-all
    (``(K/G) x (n_c/G) x 2w`` bytes per pair; C4 fp64: 16 MB) and the counts are summed with one
    all-reduce.  Bytes a rank sends per decision + PCS: ``(G-1)/G`` of
    ``3 x 8 n_c + 64 + 24`` (decision) and of ``2 w n_c K/G`` (re-shard).
    """

    def __init__(self, desc: dict, context_set: Tensor, device, dtype=torch.float64, group=None):
        from .model import B200ModelListGP
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.K = len(desc["train_X"])
        self.a0, self.a1 = shard_bounds(self.K, self.world, self.rank)
        self.model = B200ModelListGP(sub_description(desc, self.a0, self.a1), device=device, dtype=dtype)
        m = self.model
        self.n_c = context_set.shape[0]
        self.ctx = context_set.to(device=m.device, dtype=m.dtype).contiguous()
        self.ws = m.private_workspace(self.n_c)
        self.ws["ctx"].copy_(self.ctx)
        dev = m.device
        self._ext_arm = torch.zeros(self.n_c, dtype=torch.int32, device=dev)
        self._ext_mean = torch.zeros(self.n_c, dtype=m.dtype, device=dev)
        self._ext_var = torch.zeros(self.n_c, dtype=m.dtype, device=dev)
        self._pack = torch.zeros(3, self.n_c, dtype=torch.float64, device=dev)
        self._pack_all = torch.zeros(self.world, 3, self.n_c, dtype=torch.float64, device=dev)
        self._dec_all = torch.zeros(self.world, 64, dtype=torch.uint8, device=dev)
        self._psi = torch.zeros(3, dtype=torch.float64, device=dev)  # psi_best, psi_rest, max status
        self._bad = torch.zeros(1, dtype=torch.int32, device=dev)
        self._dec_host = torch.zeros(64, dtype=torch.uint8).pin_memory()
        self._dec_view = _lib.Decision.from_address(self._dec_host.data_ptr())
        # context re-shard for the PCS (equal splits: padded so that every rank sends the same row count)
        self._rows = (self.n_c + self.world - 1) // self.world
        self._kmax = max(e - b for b, e in (shard_bounds(self.K, self.world, r) for r in range(self.world)))
        self._cdf_ws = None
        self._stats = torch.zeros(8, dtype=torch.int64, device=dev)

    def launch(self, p_mode: int = _lib.CRS_P_COUNT, kernel_scale: float = 2.0, prepare: bool = False) -> None:
        """Enqueue one whole decision on the current stream; ``ws["decision"]`` holds the result."""
        m, ws, lib = self.model, self.ws, _lib.get()
        n_c, Kl, dt = self.n_c, m.num_arms, _lib.DTYPES[m.dtype]
        with torch.cuda.device(m.device):
            stream = _lib.stream_ptr(m.device)
            if prepare:
                m.prepare(0, Kl, relearn=False)
            st = C.byref(m.state)
            mean_p, var_p = ws["mean"].data_ptr(), ws["var"].data_ptr()
            _lib.check(lib.crs_posterior_grid(st, ws["ctx"].data_ptr(), n_c, 0, Kl, mean_p, var_p, Kl, 0, stream),
                       "crs_posterior_grid")
            # 1. local per-context best
            _lib.check(lib.crs_ocba_scan(mean_p, var_p, n_c, Kl, Kl, dt, self.a0, None, None, None,
                                         ws["best_arm"].data_ptr(), None, None, None, None, None, stream), "crs_ocba_scan")
            _lib.check(lib.crs_arm_pack_best(mean_p, var_p, n_c, Kl, dt, self.a0, ws["best_arm"].data_ptr(),
                                             self._pack.data_ptr(), stream), "crs_arm_pack_best")
            # 2. global best per context
            if self.world > 1:
                dist.all_gather_into_tensor(self._pack_all, self._pack, group=self.group)
                pack_all = self._pack_all
            else:
                pack_all = self._pack
            _lib.check(lib.crs_arm_combine_best(pack_all.data_ptr(), self.world, n_c, dt, self._ext_arm.data_ptr(),
                                                self._ext_mean.data_ptr(), self._ext_var.data_ptr(), stream),
                       "crs_arm_combine_best")
            # 3. local zeta minimiser against the global best, folded over ranks
            _lib.check(lib.crs_ocba_scan(mean_p, var_p, n_c, Kl, Kl, dt, self.a0, self._ext_arm.data_ptr(),
                                         self._ext_mean.data_ptr(), self._ext_var.data_ptr(), None,
                                         ws["min_zeta"].data_ptr(), ws["min_arm"].data_ptr(), ws["tie_cnt"].data_ptr(),
                                         ws["decision"].data_ptr(), ws["scan_ws"].data_ptr(), stream), "crs_ocba_scan")
            _lib.check(lib.crs_arm_record_mean(ws["decision"].data_ptr(), mean_p, Kl, dt, self.a0, stream),
                       "crs_arm_record_mean")
            if self.world > 1:
                dist.all_gather_into_tensor(self._dec_all, ws["decision"], group=self.group)
                recs = self._dec_all
            else:
                self._dec_all[0].copy_(ws["decision"])
                recs = self._dec_all
            _lib.check(lib.crs_arm_fold(recs.data_ptr(), self.world, self._ext_arm.data_ptr(), ws["decision"].data_ptr(),
                                        stream), "crs_arm_fold")
            # 4. psi partial sums at the selected context, summed over ranks
            _lib.check(lib.crs_ocba_select(st, ws["ctx"].data_ptr(), mean_p, var_p, Kl, self.a0, p_mode,
                                           float(kernel_scale), ws["decision"].data_ptr(), stream), "crs_ocba_select")
            dview = ws["decision"].view(torch.float64)
            self._psi[:2].copy_(dview[1:3])
            self._psi[2] = ws["decision"].view(torch.int32)[12].to(torch.float64)  # reserved[0]: first bad arm + 1
            if self.world > 1:
                dist.all_reduce(self._psi[:2], group=self.group)
                dist.all_reduce(self._psi[2:], op=dist.ReduceOp.MAX, group=self.group)
            self._bad.copy_(self._psi[2:].to(torch.int32))
            _lib.check(lib.crs_arm_finish(ws["decision"].data_ptr(), self._psi.data_ptr(), self._bad.data_ptr(), stream),
                       "crs_arm_finish")

    def decide_struct(self, p_mode: int = _lib.CRS_P_COUNT, kernel_scale: float = 2.0, prepare: bool = False):
        self.launch(p_mode, kernel_scale, prepare)
        self._dec_host.copy_(self.ws["decision"], non_blocking=True)
        torch.cuda.current_stream(self.model.device).synchronize()
        d = self._dec_view
        if d.reserved[0] > 0:
            raise RuntimeError(f"arm {d.reserved[0] - 1}: kernel matrix not positive definite")
        return d

    def decide(self, p_mode: int = _lib.CRS_P_COUNT, kernel_scale: float = 2.0, prepare: bool = False) -> Tensor:
        """Returns the combined record (``REC_*`` columns) on the host, identical on every rank."""
        d = self.decide_struct(p_mode, kernel_scale, prepare)
        return decision_record(d.min_zeta, d.tie_count, d.context, d.zeta_arm, d.best_arm, d.next_arm,
                               d.psi_best, d.psi_rest)

    def reshard_by_context(self) -> Tuple[Tensor, Tensor, int, int]:
        """One all-to-all of the local grid: returns ``(mean, var)`` ``[n_loc, K]`` of this rank's context
        slice with ALL arms, plus the slice's ``(begin, n_loc)``."""
        m, ws, G = self.model, self.ws, self.world
        b, e = shard_bounds(self.n_c, G, self.rank)
        if G == 1:
            return ws["mean"], ws["var"], 0, self.n_c
        Kl = m.num_arms
        even_rows = self.n_c % G == 0
        even_arms = self.K % G == 0
        out = []
        for tab in (ws["mean"], ws["var"]):
            if even_rows and even_arms:  # one contiguous exchange + one permuting copy
                recv = torch.empty(G, e - b, Kl, dtype=m.dtype, device=m.device)
                dist.all_to_all_single(recv, tab.view(G, e - b, Kl), group=self.group)
                out.append(recv.permute(1, 0, 2).reshape(e - b, self.K))
            else:
                splits = [shard_bounds(self.n_c, G, r) for r in range(G)]
                arms = [shard_bounds(self.K, G, r) for r in range(G)]
                send = [tab[s0:s1].contiguous() for s0, s1 in splits]
                recv = [torch.empty(e - b, a1 - a0, dtype=m.dtype, device=m.device) for a0, a1 in arms]
                dist.all_to_all(recv, send, group=self.group)
                out.append(torch.cat(recv, dim=1).contiguous())
        return out[0], out[1], b, e - b

    def pcs_total_launch(self, num_samples: int, seed: int, offset: int = 0) -> Tensor:
        """Context re-shard + local counts + one all-reduce; returns the 1-element int64 device tensor
        holding the total success count over all contexts (no host sync)."""
        from .generalized_pcs import pcs_counts
        mean, var, b, n_loc = self.reshard_by_context()
        if n_loc:
            if self._cdf_ws is None:
                nbytes = int(_lib.get().crs_pcs_cdf_workspace_bytes(max(n_loc, 1)))
                self._cdf_ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.model.device)
            pcs_counts(self.model, mean, var, num_samples, seed, offset, ctx_offset=b, stats=self._stats,
                       sampler="cdf", workspace=self._cdf_ws)
        else:
            self._stats.zero_()
        total = self._stats[4:5]
        if self.world > 1:
            dist.all_reduce(total, group=self.group)
        return total

    def pcs_mean(self, num_samples: int, seed: int, offset: int = 0) -> Tensor:
        total = self.pcs_total_launch(num_samples, seed, offset)
        return total.to(torch.float64).cpu() / (float(num_samples) * self.n_c)


class ShardedSequentialRS:
    """The sequential R&S loop of ``contextual_rs_b200.rs_loop.SequentialRS`` with the context set
    sharded over the ranks (BASELINE config 5: decision + PCS per iteration at 1/2/4/8 GPUs).

    Every rank keeps the whole GP state (replicated: each observation is applied by every rank —
    one arm's ``crs_gp_prepare`` + one column of the LOCAL grid) and the posterior grid of its own
    context slice.  Exchanges per iteration: one 64-byte-per-rank all-gather (decision) and one
    scalar all-reduce (PCS, ``rho`` = mean) or one all-gather of the int32 counts (general ``rho``).
    Tie randomisation is not supported across ranks (lowest context wins, like
    ``ContextShardedDecider``).
    """

    def __init__(self, model, context_set: Tensor, infer_p: bool = False, use_kde: bool = False,
                 kernel_scale: float = 1.0, num_pcs_samples: int = 64, pcs_seed: int = 0, rho=None,
                 group=None, exchange: str = "mailbox"):
        from .contextual_rs_strategies import _p_mode
        self.model = model
        self.group = group
        self.context_host = context_set.detach().cpu()
        self.dec = ContextShardedDecider(model, context_set, group=group, exchange=exchange)
        self.world, self.rank, self.n_c = self.dec.world, self.dec.rank, self.dec.n_c
        self.n_loc = self.dec.end - self.dec.begin
        self.p_mode = _p_mode(infer_p, use_kde)
        self.kernel_scale = kernel_scale
        self.num_pcs_samples, self.pcs_seed, self.rho = num_pcs_samples, pcs_seed, rho
        self.iteration = 0
        self._refresh(0, model.num_arms)

    def _refresh(self, a0: int, a1: int) -> None:
        if self.n_loc:
            ws = self.dec.ws
            self.model.posterior_tables(self.dec.local_ctx, ws["mean"], ws["var"], arms=(a0, a1))

    def decide(self) -> dict:
        """Scan + step 2(b) on the cached local grid, then the 64-byte all-gather."""
        from .contextual_rs_strategies import _scan_select
        m, d = self.model, self.dec
        if d.exchange == "mailbox":  # device-side combine, result polled from mapped host memory
            dec = d.decide_struct(self.p_mode, self.kernel_scale, None, grid=False)
            c = int(dec.context)
            return {"next_arm": int(dec.next_arm), "context_index": c, "next_context": self.context_host[c],
                    "min_zeta": float(dec.min_zeta)}
        if self.n_loc:
            ws = d.ws
            with torch.cuda.device(m.device):
                _scan_select(_lib.get(), m, ws, self.n_loc, self.p_mode, self.kernel_scale,
                             _lib.stream_ptr(m.device), ctx_ptr=d.local_ctx.data_ptr())
            local = ws["decision"]
        else:
            local = d._empty
        if self.world > 1:
            dist.all_gather_into_tensor(d._gather, local, group=self.group)
            d._gather_host.copy_(d._gather, non_blocking=True)
        else:
            d._gather_host[0].copy_(local, non_blocking=True)
        torch.cuda.current_stream(m.device).synchronize()
        rec = combine_context_sharded(records_from_bytes(d._gather_host, self.n_c, self.world))
        c = int(rec[REC_CONTEXT])
        return {"next_arm": int(rec[REC_NEXT_ARM]), "context_index": c, "next_context": self.context_host[c],
                "min_zeta": float(rec[REC_MIN_ZETA])}

    def pcs(self) -> Tensor:
        if self.rho is None and self.dec.exchange == "mailbox":  # mean over contexts: one peer sum on the device
            # (waited for here: leaving it pending — `pcs_total_async` — did not pay in this loop, whose next
            #  host call is a synchronous upload of the new observation: measured 2.27 -> 3.35 ms per iteration at N = 2)
            total = self.dec.pcs_total(self.num_pcs_samples, self.pcs_seed, offset=self.iteration)
            return torch.tensor(total / (float(self.num_pcs_samples) * self.n_c), dtype=torch.float64)
        counts = self.dec.pcs_counts(self.num_pcs_samples, self.pcs_seed, offset=self.iteration)
        if self.rho is None:  # mean over contexts: one scalar all-reduce
            total = counts.sum(dtype=torch.int64).to(torch.float64).reshape(1)
            if self.world > 1:
                dist.all_reduce(total, group=self.group)
            return (total / (float(self.num_pcs_samples) * self.n_c)).squeeze(0)
        if self.world > 1:  # general rho: gather the per-context counts (padded to equal shards)
            width = (self.n_c + self.world - 1) // self.world
            pad = torch.full((width,), -1, dtype=torch.int32, device=counts.device)
            pad[: self.n_loc] = counts[: self.n_loc]
            allc = torch.empty(self.world, width, dtype=torch.int32, device=counts.device)
            dist.all_gather_into_tensor(allc, pad, group=self.group)
            counts = allc[allc >= 0]
        pcs_c = (counts.to(self.model.dtype) / float(self.num_pcs_samples)).unsqueeze(-1)
        return self.rho(pcs_c).squeeze(-1)

    def observe(self, arm: int, x: Tensor, y: Tensor) -> None:
        m = self.model
        cap_before = m.n_cap
        m.condition_on_observations(int(arm), x, y)
        if m.n_cap != cap_before:
            self._refresh(0, m.num_arms)
        else:
            self._refresh(int(arm), int(arm) + 1)

    def step(self, evaluate) -> dict:
        """decide -> simulate -> PCS -> update (``experiments/wsc_experiments/main.py:392-468``);
        ``evaluate`` must return the same value on every rank."""
        out = self.decide()
        y = evaluate(out["next_arm"], out["next_context"])
        out["pcs"] = self.pcs()
        self.observe(out["next_arm"], out["next_context"].view(1, -1), torch.as_tensor(y).reshape(1, 1))
        self.iteration += 1
        return out
