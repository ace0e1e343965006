"""Batch-sharded (data-parallel) helpers: the path's only cross-rank exchange.

Every quantity on the path is per pixel or per sample (FiLM pools inside a sample, CE/Dice are
per-sample means, the confusion matrix is a pixel count), so the batch shards over ranks with
no pixel data ever leaving its GPU.  One process per GPU (torch.distributed, NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests); per step ONE all-reduce of a small
packed fp64 buffer turns the per-rank step summaries into the numbers the single-process
reference would report on the concatenated batch, and gives the factors that make the
DDP-averaged gradients equal to the single-process gradients.

Reference semantics being reproduced: CE is a mean over ALL samples of the
global batch; Dice is a mean over the samples whose dice is not NaN (Metrics/losses.py:64-66),
consistency a mean over B*H*W; confusion matrices add.
"""
from typing import List, Sequence

import torch
import torch.distributed as dist


def pack_step_summary(scalars: torch.Tensor, batch_local: int, confusion: Sequence[torch.Tensor]) -> torch.Tensor:
    """fp64 buffer: [B_local, consistency*B_local, per level (ce*B_local, dice*n_dice, n_dice, n_ce)] + confusion counts.
    `scalars` is StepOutput.scalars ([2 + 4*n]: total, consistency, then ce, dice, n_dice, n_ce per level).
    int64 counts are exact in fp64 up to 2^53."""
    n = (scalars.numel() - 2) // 4
    s = scalars.double()
    parts = [torch.tensor([float(batch_local)], dtype=torch.float64, device=s.device), (s[1] * batch_local).reshape(1)]
    for L in range(n):
        ce, dice, n_dice, n_ce = s[2 + 4 * L], s[3 + 4 * L], s[4 + 4 * L], s[5 + 4 * L]
        parts.append(torch.stack([ce * batch_local, dice * n_dice, n_dice, n_ce]))
    parts += [c.reshape(-1).double() for c in confusion]
    return torch.cat(parts)


def unpack_global(buf: torch.Tensor, n_levels: int, conf_shapes: Sequence[Sequence[int]]):
    """Inverse of pack_step_summary after the SUM all-reduce: global loss terms + confusion matrices."""
    B = buf[0]
    out = {"batch": B, "consistency": buf[1] / B, "ce": [], "dice": [], "n_dice": [], "n_ce": []}
    off = 2
    for _ in range(n_levels):
        ce_sum, dice_sum, n_dice, n_ce = buf[off:off + 4]
        out["ce"].append(ce_sum / B)
        out["dice"].append(torch.where(n_dice > 0, dice_sum / torch.clamp(n_dice, min=1.0), torch.zeros_like(dice_sum)))
        out["n_dice"].append(n_dice)
        out["n_ce"].append(n_ce)
        off += 4
    conf = []
    for shp in conf_shapes:
        k = int(shp[0]) * int(shp[1])
        conf.append(buf[off:off + k].round().to(torch.int64).view(int(shp[0]), int(shp[1])))
        off += k
    out["confusion"] = conf
    out["total"] = sum(out["ce"]) + sum(out["dice"]) + out["consistency"]
    return out


def all_reduce_step_summary(scalars, batch_local, confusion, group=None):
    """The step's single collective.  Returns the unpacked global summary (see unpack_global)."""
    buf = pack_step_summary(scalars, batch_local, confusion)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    n = (scalars.numel() - 2) // 4
    return unpack_global(buf, n, [tuple(c.shape) for c in confusion])


def pack_exchange(summary: torch.Tensor, extra: Sequence[torch.Tensor] = (), out: torch.Tensor = None) -> torch.Tensor:
    """The step's exchange buffer [summary | extra tensors] in fp64.  CUDA: the fp32 `extra` tensors are
    written behind the summary by ONE kernel (rhseg_pack_f64); with `out` = StepOutput.exchange (allocated by a
    FusedHierStep whose exchange_tail covers the extras) the summary is already in place and nothing is copied.
    CPU tensors (gloo tests of the host logic) are concatenated with torch."""
    if not summary.is_cuda:
        flat = [summary] + [e.reshape(-1).double() for e in extra]
        return torch.cat(flat) if len(flat) > 1 else summary.clone()
    import ctypes
    from . import native
    extra = [e if (e.dtype == torch.float32 and e.is_contiguous()) else e.float().contiguous() for e in extra]
    n_sum = summary.numel()
    need = n_sum + sum(e.numel() for e in extra)
    if out is not None and out.data_ptr() == summary.data_ptr() and out.numel() >= need and out.dtype == torch.float64:
        buf = out[:need]
    else:
        buf = torch.empty(need, dtype=torch.float64, device=summary.device)
        buf[:n_sum].copy_(summary)
    if extra:
        ptrs = (ctypes.c_void_p * len(extra))(*[e.data_ptr() for e in extra])
        cnts = (ctypes.c_long * len(extra))(*[e.numel() for e in extra])
        native.call("rhseg_pack_f64", ptrs, cnts, len(extra), buf.data_ptr() + 8 * n_sum, native.stream_of(summary))
    return buf


def unpack_exchange(buf: torch.Tensor, n_summary: int, targets: Sequence[torch.Tensor], scale: float = 1.0) -> None:
    """Inverse of pack_exchange for the gradient part after the all-reduce: targets[k] <- scale * buf slice
    (fp32, in place; scale = 1/world gives DDP's average).  CUDA: one rhseg_unpack_f32 launch."""
    if not buf.is_cuda:
        off = n_summary
        for t in targets:
            t.copy_((buf[off:off + t.numel()] * scale).view(t.shape).to(t.dtype))
            off += t.numel()
        return
    import ctypes
    from . import native
    for t in targets:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise native.NativeError("unpack_exchange writes contiguous float32 tensors in place")
    if targets:
        ptrs = (ctypes.c_void_p * len(targets))(*[t.data_ptr() for t in targets])
        cnts = (ctypes.c_long * len(targets))(*[t.numel() for t in targets])
        native.call("rhseg_unpack_f32", buf.data_ptr() + 8 * n_summary, float(scale), ptrs, cnts, len(targets),
                    native.stream_of(buf))


class PeerExchange:
    """One-shot all-reduce of the exchange buffer over NVLink peer memory (rhseg_xchg_*): single node, one
    process per GPU.  Construct it once, collectively, after init_process_group (every rank of `group`):

        px = PeerExchange(capacity=out.exchange.numel())
        buf = px.all_reduce(out.summary, param_grads, out=out.exchange)     # one kernel, graph-capturable

    Raises NativeError when peer mapping is impossible (no P2P between the GPUs); callers then keep the
    NCCL all-reduce of pack_exchange().  Multi-node jobs use NCCL."""

    def __init__(self, capacity: int, group=None):
        import ctypes
        from . import native
        if not (dist.is_available() and dist.is_initialized()):
            raise native.NativeError("PeerExchange needs an initialised torch.distributed process group")
        if not torch.cuda.is_available():
            raise native.NativeError("PeerExchange needs CUDA devices (peer memory over NVLink); there is no CPU form")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.capacity = int(capacity)
        self._ctx = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * native.XCHG_HANDLE_BYTES)()
        dev = torch.device("cuda", torch.cuda.current_device())

        def all_ok(rc):  # collective: every rank learns whether every rank succeeded (nobody is left waiting)
            ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            return int(ok.item()) == 1

        def fail(stage, rc):
            if self._ctx:
                native.lib().rhseg_xchg_destroy(self._ctx)
            self._ctx = None
            raise native.NativeError("%s failed on at least one rank (local status %d: %s)" % (stage, rc, native.status_string(rc)))

        rc = native.lib().rhseg_xchg_create(self.capacity, self.world, ctypes.byref(self._ctx), handle)
        if rc != 0:
            self._ctx = None
        if not all_ok(rc):
            fail("rhseg_xchg_create", rc)
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=group)
        blob = bytes(torch.cat(gathered).cpu().tolist())
        rc = native.lib().rhseg_xchg_connect(self._ctx, self.rank, blob)
        if not all_ok(rc):  # every rank must have mapped its peers before anybody uses the buffers
            fail("rhseg_xchg_connect", rc)

    def all_reduce(self, summary: torch.Tensor, extra: Sequence[torch.Tensor] = (), out: torch.Tensor = None) -> torch.Tensor:
        """SUM over ranks of [summary | extra...] (fp64), written to `out` (default: a new tensor; `out` may be the
        StepOutput.exchange buffer whose prefix is `summary`)."""
        import ctypes
        from . import native
        native.require_cuda(summary, *extra)
        if summary.dtype != torch.float64 or not summary.is_contiguous():
            raise native.NativeError("the step summary is a contiguous float64 tensor")
        extra = [e if (e.dtype == torch.float32 and e.is_contiguous()) else e.float().contiguous() for e in extra]
        need = summary.numel() + sum(e.numel() for e in extra)
        if need > self.capacity:
            raise native.NativeError("exchange of %d elements exceeds the PeerExchange capacity %d" % (need, self.capacity))
        if out is None or out.numel() < need or out.dtype != torch.float64:
            out = torch.empty(need, dtype=torch.float64, device=summary.device)
        ptrs = (ctypes.c_void_p * max(1, len(extra)))(*[e.data_ptr() for e in extra])
        cnts = (ctypes.c_long * max(1, len(extra)))(*[e.numel() for e in extra])
        native.call("rhseg_xchg_all_reduce", self._ctx, summary.data_ptr(), summary.numel(), ptrs, cnts, len(extra),
                    out.data_ptr(), native.stream_of(summary))
        return out[:need]

    def status(self) -> int:
        """0, or 1 (sticky) when a wait for a peer timed out (synchronises the device)."""
        import ctypes
        from . import native
        s = ctypes.c_int(0)
        native.check(native.lib().rhseg_xchg_status(self._ctx, ctypes.byref(s)), "rhseg_xchg_status")
        return int(s.value)

    def check(self) -> None:
        """Raises when an exchange timed out.  A timeout is a hard failure: the kernel poisons that result and every later
        one with NaN (so loss and gradients turn NaN rather than the replicas diverging silently); call this at a host
        sync point the training loop already has (where it reads the loss, or around the optimiser step)."""
        from . import native
        if self.status() != 0:
            raise native.NativeError("peer-memory exchange timed out waiting for a rank (rank %d of %d): results since then "
                                     "are NaN; the job must stop" % (self.rank, self.world))

    def set_timeout_ms(self, ms: int) -> None:
        """Wait limit per exchange (default 30 s, RHSEG_XCHG_TIMEOUT_MS); 0 waits for ever, like an NCCL all-reduce.
        Applies to exchanges launched (or captured into a graph) afterwards."""
        from . import native
        native.check(native.lib().rhseg_xchg_set_timeout_ms(self._ctx, int(ms)), "rhseg_xchg_set_timeout_ms")

    def close(self):
        if getattr(self, "_ctx", None):
            from . import native
            torch.cuda.synchronize()
            native.lib().rhseg_xchg_destroy(self._ctx)
            self._ctx = None


def all_reduce_summary(summary: torch.Tensor, n_levels: int, conf_shapes, extra: Sequence[torch.Tensor] = (), group=None):
    """Fused-step variant: `summary` is StepOutput.summary (already in the packed layout, written by
    rhseg_step_finalize).  `extra` tensors (e.g. the head / FiLM parameter gradients) ride in the same
    all-reduce.  Returns (global summary dict, reduced extras as fp64 views)."""
    buf = pack_exchange(summary, extra)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out = unpack_global(buf, n_levels, conf_shapes)
    views, off = [], summary.numel()
    for e in extra:
        views.append(buf[off:off + e.numel()].view(e.shape))
        off += e.numel()
    return out, views


def dice_grad_scale(n_dice_local: torch.Tensor, n_dice_global: torch.Tensor, world_size: int) -> torch.Tensor:
    """Factor for a rank's Dice gradient so that DDP's mean over ranks equals the gradient of the
    single-process Dice (a mean over the GLOBAL count of valid samples): world * n_local / n_global.
    CE needs no correction when every rank holds the same number of samples."""
    return torch.where(n_dice_global > 0, world_size * n_dice_local / torch.clamp(n_dice_global, min=1.0),
                       torch.zeros_like(n_dice_global))


def shard_batch(batch: int, rank: int, world: int):
    """Contiguous batch slice [start, stop) of `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
