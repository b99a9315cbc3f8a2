"""Drop-ins for the forecast metrics of `ladcast.evaluate.utils` (reference evaluate/utils.py:40-118) and the
per-lead-time assembly of `evaluate/evaluate_ens_gpu.py:339-415`, computed by the on-device reduction kernel
(`lc_metrics_*`).  With members sharded over GPUs, `ensemble_metrics_distributed` re-shards the decoded fields
(member-major -> (channel, lead)-plane-major) with one NCCL all-to-all, reduces locally and all-gathers the tables."""
from typing import Callable, Dict, Optional, Union

import numpy as np
import torch

from .. import _lib

SST_CHANNEL_IDX = 82
METRIC_NAMES = ("ens_mse", "crps_skill", "crps_spread", "crps")


def get_normalized_lat_weights_based_on_cos(lat: Union[torch.Tensor, np.ndarray]):
    """requires lat in degrees; cos(lat) / mean(cos(lat)) in the input's precision (numpy float64 in the scripts)."""
    if isinstance(lat, torch.Tensor):
        w = torch.cos(torch.deg2rad(lat))
    else:
        w = np.cos(np.deg2rad(lat))
    return w / w.mean()


def _planes(forecast: torch.Tensor, ensemble_dim: int):
    f = forecast.movedim(ensemble_dim, 0).to(torch.float32).contiguous()
    M, H, W = f.shape[0], f.shape[-2], f.shape[-1]
    return f, M, int(np.prod(f.shape[1:-2])) if f.dim() > 3 else 1, H, W


def _pointwise(forecast, truth, ensemble_dim, want):
    if not forecast.is_cuda:
        raise _lib.LadcastB200Error("metrics run on CUDA tensors only; there is no CPU fallback")
    f, M, N, H, W = _planes(forecast, ensemble_dim)
    out_shape = f.shape[1:]
    t = None
    if truth is not None:
        t = torch.broadcast_to(truth.movedim(ensemble_dim, 0) if truth.dim() == forecast.dim() else truth.unsqueeze(0),
                               f.shape[:1] + out_shape)[0].to(torch.float32).contiguous()
    res = {k: torch.empty(out_shape, device=f.device, dtype=torch.float32) for k in want}
    lib = _lib.load()
    with torch.cuda.device(f.device):
        _lib.check(lib.lc_metrics_pointwise(_lib.ptr(f), _lib.ptr(t), M, N, H, W, _lib.ptr(res.get("skill")),
                                            _lib.ptr(res.get("spread")), _lib.ptr(res.get("mean")), _lib.stream()),
                   "lc_metrics_pointwise")
    return res


@torch.no_grad()
def pointwise_crps_skill(forecast: torch.Tensor, truth: torch.Tensor, ensemble_dim: int) -> torch.Tensor:
    """mean_m |truth - forecast_m| (truth broadcastable to forecast)."""
    return _pointwise(forecast, truth, ensemble_dim, ("skill",))["skill"]


@torch.no_grad()
def pointwise_crps_spread(forecast: torch.Tensor, ensemble_dim: int) -> torch.Tensor:
    """2/(M(M-1)) sum_i (2i - M - 1) x_(i)  ==  mean absolute difference over member pairs; zeros for M < 2."""
    return _pointwise(forecast, None, ensemble_dim, ("spread",))["spread"]


@torch.no_grad()
def get_crps(forecast: torch.Tensor, truth: torch.Tensor, ensemble_dim: int = 0) -> torch.Tensor:
    r = _pointwise(forecast, truth, ensemble_dim, ("skill", "spread"))
    return r["skill"] - 0.5 * r["spread"]


@torch.no_grad()
def ensemble_mean(forecast: torch.Tensor, ensemble_dim: int = 0) -> torch.Tensor:
    """Mean over the ensemble dimension on the device (the pointwise mode of the metrics kernel) — what
    roll_out_serial(return_ensemble_mean=True) keeps of every decoded block (pipelines/utils.py:608-630)."""
    return _pointwise(forecast, None, ensemble_dim, ("mean",))["mean"]


@torch.no_grad()
def get_acc(forecast: torch.Tensor, truth: torch.Tensor, climate: torch.Tensor,
            lat_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Anomaly correlation coefficient over the last two (spatial) dims with nanmean semantics (reference
    evaluate/utils.py:122-149); forecast / truth / climate [..., H, W] broadcastable, lat_weight broadcastable [H, 1]."""
    shape = torch.broadcast_shapes(forecast.shape, truth.shape, climate.shape)
    H, W = shape[-2], shape[-1]
    dev = forecast.device
    f, t, c = [torch.broadcast_to(x.to(dev, torch.float32), shape).reshape(-1, H * W).contiguous() for x in (forecast, truth, climate)]
    N = f.shape[0]
    lw = None
    if lat_weight is not None:
        lw = torch.broadcast_to(torch.as_tensor(lat_weight).to(dev, torch.float64).reshape(-1, 1) if torch.as_tensor(lat_weight).numel() == H
                                else torch.as_tensor(lat_weight).to(dev, torch.float64), (H, 1)).reshape(H).contiguous()
    lib = _lib.load()
    sm = torch.empty((3, N), device=dev, dtype=torch.float64)
    cnt = torch.empty((3, N), device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        _lib.check(lib.lc_metrics_acc(_lib.ptr(f), _lib.ptr(t), _lib.ptr(c), _lib.ptr(lw), N, H, W, _lib.ptr(sm), _lib.ptr(cnt),
                                      _lib.stream()), "lc_metrics_acc")
    m = sm / cnt  # nanmean of the three weighted products
    out = m[0] / torch.sqrt(m[1] * m[2])
    res = out.reshape(shape[:-2])
    return res if lat_weight is not None and torch.as_tensor(lat_weight).dtype == torch.float64 else res.to(torch.float32)


def climatology_to_timeseries(clim, start_time, lead_time, interval=6, exclude_start=True, hours=(0, 6, 12, 18)):
    """Array form of the reference's `climatology_to_timeseries` (evaluate/utils.py:152-201; xarray is not in this image):
    `clim` [366 (dayofyear 1..366), len(hours), ...] (tensor or ndarray) -> ([nt, ...] climatology along the forecast
    valid times start_time (+ interval) ... start_time + lead_time h, list of datetimes).  The selection is by
    (dayofyear, hour) LABEL exactly like `ds.sel(dayofyear=..., hour=...)`: an hour that is not in `hours` raises."""
    from datetime import datetime, timedelta

    if clim.ndim < 2:
        raise ValueError("Dataset must have both 'dayofyear' and 'hour' dims")
    if isinstance(start_time, np.datetime64):
        start = start_time.astype("datetime64[s]").astype(datetime)
    else:
        start = start_time if isinstance(start_time, datetime) else datetime.fromisoformat(str(start_time))
    n = int(lead_time) // int(interval)
    times = [start + timedelta(hours=int(interval) * k) for k in range(n + 1)]
    if exclude_start:
        times = times[1:]
    hour_index = {int(h): i for i, h in enumerate(hours)}
    try:
        hi = [hour_index[t.hour] for t in times]
    except KeyError as e:
        raise KeyError(f"hour {e.args[0]} is not a climatology hour {tuple(hours)}") from None
    di = [t.timetuple().tm_yday - 1 for t in times]
    if isinstance(clim, torch.Tensor):
        idx = (torch.as_tensor(di, device=clim.device), torch.as_tensor(hi, device=clim.device))
        return clim[idx], times
    return np.asarray(clim)[np.asarray(di), np.asarray(hi)], times


def _tables_from_sums(sums, counts, n_pix, channels, leads, sst_channel):
    """sums/counts [4, C*T] fp64 -> dict of [C, T] fp64 tables: mean over pixels, nanmean for the SST channel, NaN
    propagation elsewhere (torch.mean semantics)."""
    out = {}
    for k, name in enumerate(METRIC_NAMES):
        s, c = sums[k].reshape(channels, leads), counts[k].reshape(channels, leads)
        tab = s / n_pix
        tab = torch.where(c < n_pix, torch.full_like(tab, float("nan")), tab)
        if sst_channel is not None and 0 <= sst_channel < channels:
            tab[sst_channel] = s[sst_channel] / c[sst_channel]
        out[name] = tab
    return out


def _local_sums_cuda(fields: torch.Tensor, truth: torch.Tensor, lat_weights: torch.Tensor):
    """fields [M, N, H, W] f32 (planes of a member contiguous, any member stride), truth [N, H, W] f32 ->
    (sums [4, N], counts [4, N]) fp64 via the CUDA kernel, read in place (no staging copies)."""
    lib = _lib.load()
    M, N, H, W = fields.shape
    if fields.dtype != torch.float32 or (M > 1 and fields[0].stride() != (H * W, W, 1)) or fields.stride()[1:] != (H * W, W, 1):
        fields = fields.to(torch.float32).contiguous()
    truth = truth.to(fields.device, torch.float32).contiguous()
    sums = torch.empty((4, N), device=fields.device, dtype=torch.float64)
    counts = torch.empty((4, N), device=fields.device, dtype=torch.float64)
    lw = lat_weights.to(fields.device, torch.float64).contiguous()
    with torch.cuda.device(fields.device):
        _lib.check(lib.lc_metrics_accumulate_strided(_lib.ptr_any(fields), fields.stride(0) if M > 1 else N * H * W,
                                                     _lib.ptr(truth), _lib.ptr(lw), M, N, H, W, _lib.ptr(sums),
                                                     _lib.ptr(counts), _lib.stream()), "lc_metrics_accumulate_strided")
    return sums, counts


@torch.no_grad()
def ensemble_metrics(fields: torch.Tensor, truth: torch.Tensor, lat_weights=None,
                     sst_channel: Optional[int] = SST_CHANNEL_IDX) -> Dict[str, torch.Tensor]:
    """fields [M, C, T, H, W], truth [C, T, H, W] (NaN allowed) -> {ens_mse, crps_skill, crps_spread, crps}: [C, T] fp64
    tables, exactly the quantities evaluate_ens_gpu.py stores per lead time (RMSE = sqrt(ens_mse) downstream)."""
    M, C, T, H, W = fields.shape
    if lat_weights is None:
        lat_weights = torch.from_numpy(get_normalized_lat_weights_based_on_cos(np.linspace(-88.5, 90, H)))
    if not fields.is_cuda:
        raise _lib.LadcastB200Error("metrics run on CUDA tensors only; there is no CPU fallback")
    f = fields.to(torch.float32)
    f = f.reshape(M, C * T, H, W) if f[0].is_contiguous() else f.contiguous().reshape(M, C * T, H, W)
    sums, counts = _local_sums_cuda(f, truth.to(fields.device, torch.float32).reshape(C * T, H, W), torch.as_tensor(lat_weights))
    return _tables_from_sums(sums, counts, H * W, C, T, sst_channel)


def plane_shard(n_planes: int, rank: int, world: int) -> range:
    lo, hi = (rank * n_planes) // world, ((rank + 1) * n_planes) // world
    return range(lo, hi)


def exchange_bytes(members, n_planes: int, hw: int, rank: int, world: int) -> int:
    """fp32 bytes rank `rank` SENDS to other ranks in the member->plane re-shard (its members' values on foreign planes)."""
    mine = len(plane_shard(n_planes, rank, world))
    return 4 * hw * members[rank] * (n_planes - mine)


@torch.no_grad()
def reshard_members_to_planes(fields_local: torch.Tensor, members, group=None, out: Optional[torch.Tensor] = None):
    """The one exchange of the path: fields_local [M_r, N, HW] (this rank's members, all planes) ->
    [sum(members), n_mine, HW] (all members, this rank's contiguous slice of the N planes).  One grouped NCCL
    send/recv (ncclGroupStart .. ncclSend/ncclRecv per (member, peer) .. ncclGroupEnd via batch_isend_irecv): every
    message is a contiguous slice of the source and lands in its final position of the destination, so there is no
    pack / unpack copy on either side.  Works unchanged on gloo (CPU tests)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    M_r, N, HW = fields_local.shape
    mine = plane_shard(N, rank, world)
    n_mine = len(mine)
    total = sum(members)
    first = [sum(members[:q]) for q in range(world)]  # global index of rank q's first member
    if out is None:
        out = torch.empty((total, n_mine, HW), dtype=fields_local.dtype, device=fields_local.device)
    ops = []
    for q in range(world):
        pq = plane_shard(N, q, world)
        if q == rank:
            continue
        for m in range(M_r):
            if len(pq):
                ops.append(dist.P2POp(dist.isend, fields_local[m, pq.start : pq.stop], _global_rank(group, q), group))
        for l in range(members[q]):
            if n_mine:
                ops.append(dist.P2POp(dist.irecv, out[first[q] + l], _global_rank(group, q), group))
    if n_mine and M_r:
        out[first[rank] : first[rank] + M_r].copy_(fields_local[:, mine.start : mine.stop])  # my own share: local copy
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


class _DevPtrHolder:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerBuffer:
    """One fp32 device buffer of `numel` elements per rank, each mapped into every other process of the node
    (`lc_ipc_alloc` / `lc_ipc_open`: CUDA IPC opened on the reader's own GPU, peer access over NVLink / NVSwitch).
    `tensor` is this rank's buffer as a torch tensor; `ptrs[q]` is the address at which THIS process sees rank q's buffer.
    Collective constructor (handles travel through all_gather_object); keep the object alive while kernels use it."""

    def __init__(self, numel: int, device: torch.device, group=None):
        import ctypes

        import torch.distributed as dist

        self._lib = _lib.load()
        self.device, self.numel = device, int(numel)
        self._group = group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        # Every rank walks through every collective of this constructor even after a local failure, and all ranks
        # raise together: a rank that bailed out alone would leave the others hanging in the next collective.
        self._own, self._opened, self.ptrs, self.tensor = None, [], [], None
        err = None
        own, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        try:
            with torch.cuda.device(device):
                _lib.check(self._lib.lc_ipc_alloc(4 * self.numel, ctypes.byref(own), handle), "lc_ipc_alloc")
            self._own = own.value
            self._holder = _DevPtrHolder(self._own, (self.numel,))
            self.tensor = torch.as_tensor(self._holder, device=device)
        except Exception as e:  # noqa: BLE001
            err = e
        handles = [None] * world
        dist.all_gather_object(handles, None if err is not None else bytes(handle), group=group)
        if err is None and any(h is None for h in handles):
            err = _lib.LadcastB200Error("a peer rank could not allocate its peer-visible buffer")
        if err is None:
            try:
                with torch.cuda.device(device):
                    for q in range(world):
                        if q == rank:
                            self.ptrs.append(self._own)
                            continue
                        p = ctypes.c_void_p()
                        hq = (ctypes.c_ubyte * 64).from_buffer_copy(handles[q])
                        _lib.check(self._lib.lc_ipc_open(hq, ctypes.byref(p)), "lc_ipc_open")
                        self.ptrs.append(p.value)
                        self._opened.append(p.value)
            except Exception as e:  # noqa: BLE001
                err = e
        oks = [None] * world
        dist.all_gather_object(oks, err is None, group=group)  # doubles as the barrier: every mapping exists before use
        if not all(oks):
            self._abandon()
            raise err if err is not None else _lib.LadcastB200Error("a peer rank could not map the peer-visible buffers")

    def _abandon(self):
        """Local clean-up after a failed (collective) construction: no further collectives."""
        with torch.cuda.device(self.device):
            for p in self._opened:
                self._lib.lc_ipc_close(p)
            self.tensor = None
            if self._own is not None:
                self._lib.lc_ipc_free(self._own)
        self._own, self._opened, self.ptrs = None, [], []

    def release(self):
        """Collective: unmap the peers' buffers, then (after a barrier) free the own one."""
        import torch.distributed as dist

        if self._own is None:
            return
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for p in self._opened:
                self._lib.lc_ipc_close(p)
        dist.barrier(group=self._group)
        self.tensor = None
        with torch.cuda.device(self.device):
            self._lib.lc_ipc_free(self._own)
        self._own, self._opened, self.ptrs = None, [], []


_PEER_BUFFERS: dict = {}


def _peer_buffer(numel: int, device: torch.device, group=None) -> PeerBuffer:
    """Cached PeerBuffer of at least `numel` elements per rank (the same `numel` must be requested on every rank).
    Growing the buffer frees the previous one: tensors that view it (e.g. a rollout's `out=`) must be dropped first."""
    key = (device.index, id(group))
    buf = _PEER_BUFFERS.get(key)
    if buf is None or buf.numel < numel:
        if buf is not None:
            buf.release()
        buf = PeerBuffer(numel, device, group)
        _PEER_BUFFERS[key] = buf
    return buf


def release_peer_buffers():
    """Collective: free the cached peer-visible buffers (call before destroying the process group)."""
    for buf in list(_PEER_BUFFERS.values()):
        buf.release()
    _PEER_BUFFERS.clear()


@torch.no_grad()
def _local_sums_peer(buf: PeerBuffer, members, n_planes: int, mine: range, truth_mine: torch.Tensor,
                     lat_weights: torch.Tensor, H, W):
    """sums/counts [4, len(mine)] of this rank's plane slice with every member read where it lies: rank q's members
    are the leading [M_q, N, HW] fp32 block of its peer buffer (seen here at buf.ptrs[q])."""
    import ctypes

    lib = _lib.load()
    dev = truth_mine.device
    n_mine = len(mine)
    ptrs = []
    for q, base in enumerate(buf.ptrs):
        for l in range(members[q]):
            ptrs.append(base + 4 * (l * n_planes + mine.start) * H * W)
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    sums = torch.empty((4, n_mine), device=dev, dtype=torch.float64)
    counts = torch.empty((4, n_mine), device=dev, dtype=torch.float64)
    lw = lat_weights.to(dev, torch.float64).contiguous()
    with torch.cuda.device(dev):
        _lib.check(lib.lc_metrics_accumulate_ptrs(arr, _lib.ptr(truth_mine), _lib.ptr(lw), len(ptrs), n_mine, H, W,
                                                  _lib.ptr(sums), _lib.ptr(counts), _lib.stream()),
                   "lc_metrics_accumulate_ptrs")
    return sums, counts


def _global_rank(group, q):
    import torch.distributed as dist

    return q if group is None else dist.get_global_rank(group, q)


@torch.no_grad()
def ensemble_metrics_distributed(fields_local: torch.Tensor, truth: torch.Tensor, lat_weights=None, group=None,
                                 sst_channel: Optional[int] = SST_CHANNEL_IDX,
                                 local_sums_fn: Optional[Callable] = None, timings: Optional[dict] = None,
                                 exchange: str = "nccl") -> Dict[str, torch.Tensor]:
    """Members are sharded over ranks (`fields_local` [M_r, C, T, H, W] with possibly different M_r per rank); CRPS
    spread and the ensemble mean need all members per grid point, so the fields are re-sharded once — rank r receives
    every member's values for its contiguous slice of the C*T (channel, lead) planes (`reshard_members_to_planes`) —
    reduced locally by the metrics kernel, and the [8, planes] partial tables are all-gathered.  Every rank returns the
    full [C, T] tables (reference assembly: evaluate/evaluate_ens_gpu.py:339-415, gather :462-468).
    `timings` (optional dict) receives CUDA-event milliseconds of the exchange and of the local reduction.
    exchange="p2p" (one node, CUDA): no exchange step at all — the other ranks' field blocks are mapped into this
    process (`PeerBuffer`: CUDA IPC opened on the reader's GPU) and `lc_metrics_accumulate_ptrs` reads every member of
    this rank's plane slice in place over NVLink, so the transfer overlaps the reduction pixel by pixel and the gathered
    copy (a write + a read of M * planes/world * H * W * 4 B per rank) never exists.  Fields that do not already live
    in the rank's peer buffer are copied into it first (one local device copy, reported as `exchange_ms`)."""
    import torch.distributed as dist

    if local_sums_fn is None:
        local_sums_fn = _local_sums_cuda
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    M_r, C, T, H, W = fields_local.shape
    N = C * T
    if lat_weights is None:
        lat_weights = torch.from_numpy(get_normalized_lat_weights_based_on_cos(np.linspace(-88.5, 90, H)))
    lat_weights = torch.as_tensor(lat_weights)
    dev = fields_local.device
    f = fields_local.to(torch.float32)
    f = (f if f.is_contiguous() else f.contiguous()).reshape(M_r, N, H * W)
    counts_m = torch.zeros(world, dtype=torch.int64, device=dev)
    counts_m[rank] = M_r
    dist.all_reduce(counts_m, group=group)
    members = [int(v) for v in counts_m.tolist()]
    mine = plane_shard(N, rank, world)
    n_mine = len(mine)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if (timings is not None and dev.type == "cuda") else None
    t_mine = truth.to(dev, torch.float32).reshape(N, H, W)[mine.start : mine.stop].contiguous()
    if exchange == "p2p":
        if dev.type != "cuda":
            raise _lib.LadcastB200Error("exchange='p2p' needs CUDA tensors (peer memory over NVLink)")
        if sum(members) > 64:
            raise _lib.LadcastB200Error("exchange='p2p' supports up to 64 members in total")
        buf = _peer_buffer(max(members) * N * H * W, dev, group)
        if ev:
            ev[0].record()
        if M_r > 0 and f.data_ptr() != buf.tensor.data_ptr():  # fields produced elsewhere: one local device copy
            lo = buf.tensor.data_ptr()
            if lo < f.data_ptr() < lo + 4 * buf.numel:  # a view INSIDE the peer buffer but not at its start
                f = f.clone()
            buf.tensor[: f.numel()].copy_(f.reshape(-1))
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)  # every rank's fields are complete before anybody reads them
        if ev:
            ev[1].record()
        if n_mine:
            sums_l, counts_l = _local_sums_peer(buf, members, N, mine, t_mine, lat_weights, H, W)
        else:
            sums_l = counts_l = torch.zeros((4, 0), dtype=torch.float64, device=dev)
        if ev:
            ev[2].record()
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)  # peers have finished reading this rank's fields
    elif exchange == "nccl":
        if ev:
            ev[0].record()
        gathered = reshard_members_to_planes(f, members, group)
        if ev:
            ev[1].record()
        if n_mine:
            sums_l, counts_l = local_sums_fn(gathered.reshape(-1, n_mine, H, W), t_mine, lat_weights)
        else:
            sums_l = counts_l = torch.zeros((4, 0), dtype=torch.float64, device=dev)
        if ev:
            ev[2].record()
    else:
        raise ValueError(f"unknown exchange {exchange!r}")
    n_max = max(len(plane_shard(N, q, world)) for q in range(world))
    packed = torch.zeros((8, n_max), dtype=torch.float64, device=dev)
    packed[:4, :n_mine], packed[4:, :n_mine] = sums_l, counts_l
    allp = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(allp, packed, group=group)
    sums = torch.cat([allp[q][:4, : len(plane_shard(N, q, world))] for q in range(world)], dim=1)
    counts = torch.cat([allp[q][4:, : len(plane_shard(N, q, world))] for q in range(world)], dim=1)
    if ev:
        torch.cuda.synchronize(dev)
        timings["exchange_ms"] = ev[0].elapsed_time(ev[1])
        timings["kernel_ms"] = ev[1].elapsed_time(ev[2])
        timings["bytes_sent"] = exchange_bytes(members, N, H * W, rank, world)
        timings["members"] = members
    return _tables_from_sums(sums, counts, H * W, C, T, sst_channel)
