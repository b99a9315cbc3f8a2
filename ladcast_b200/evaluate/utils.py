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
    done = 0
    fv = f.reshape(M, N, H * W)
    while done < N:  # the kernel takes at most 65535 planes per launch
        n = min(N - done, 65535)
        fs = fv[:, done : done + n].contiguous() if n != N else fv
        ts = t.reshape(N, H * W)[done : done + n].contiguous() if t is not None else None
        outs = {k: torch.empty((n, H * W), device=f.device, dtype=torch.float32) for k in want}
        _lib.check(lib.lc_metrics_pointwise(_lib.ptr(fs), _lib.ptr(ts), M, n, H, W, _lib.ptr(outs.get("skill")),
                                            _lib.ptr(outs.get("spread")), _lib.ptr(outs.get("mean")), _lib.stream()),
                   "lc_metrics_pointwise")
        for k in want:
            res[k].reshape(N, H * W)[done : done + n] = outs[k]
        done += n
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
    out = torch.empty(N, device=dev, dtype=torch.float64)
    done = 0
    while done < N:
        n = min(N - done, 65535)
        s = torch.empty((3, n), device=dev, dtype=torch.float64)
        k = torch.empty((3, n), device=dev, dtype=torch.float64)
        _lib.check(lib.lc_metrics_acc(_lib.ptr(f[done : done + n]), _lib.ptr(t[done : done + n]), _lib.ptr(c[done : done + n]),
                                      _lib.ptr(lw), n, H, W, _lib.ptr(s), _lib.ptr(k), _lib.stream()), "lc_metrics_acc")
        m = s / k
        out[done : done + n] = m[0] / torch.sqrt(m[1] * m[2])
        done += n
    res = out.reshape(shape[:-2])
    return res if lat_weight is not None and torch.as_tensor(lat_weight).dtype == torch.float64 else res.to(torch.float32)


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
    """fields [M, N, H, W] f32, truth [N, H, W] f32 -> (sums [4,N], counts [4,N]) fp64 via the CUDA kernel."""
    lib = _lib.load()
    M, N, H, W = fields.shape
    sums = torch.empty((4, N), device=fields.device, dtype=torch.float64)
    counts = torch.empty((4, N), device=fields.device, dtype=torch.float64)
    lw = lat_weights.to(fields.device, torch.float64).contiguous()
    done = 0
    while done < N:
        n = min(N - done, 65535)
        fs = fields[:, done : done + n].contiguous()
        ts = truth[done : done + n].contiguous()
        s = torch.empty((4, n), device=fields.device, dtype=torch.float64)
        c = torch.empty((4, n), device=fields.device, dtype=torch.float64)
        _lib.check(lib.lc_metrics_accumulate(_lib.ptr(fs), _lib.ptr(ts), _lib.ptr(lw), M, n, H, W, _lib.ptr(s), _lib.ptr(c),
                                             _lib.stream()), "lc_metrics_accumulate")
        sums[:, done : done + n], counts[:, done : done + n] = s, c
        done += n
    return sums, counts


@torch.no_grad()
def ensemble_metrics(fields: torch.Tensor, truth: torch.Tensor, lat_weights=None,
                     sst_channel: Optional[int] = SST_CHANNEL_IDX) -> Dict[str, torch.Tensor]:
    """fields [M, C, T, H, W], truth [C, T, H, W] (NaN allowed) -> {ens_mse, crps_skill, crps_spread, crps}: [C, T] fp64
    tables, exactly the quantities evaluate_ens_gpu.py stores per lead time (RMSE = sqrt(ens_mse) downstream)."""
    M, C, T, H, W = fields.shape
    if lat_weights is None:
        lat_weights = torch.from_numpy(get_normalized_lat_weights_based_on_cos(np.linspace(-88.5, 90, H)))
    f = fields.to(torch.float32).reshape(M, C * T, H, W)
    sums, counts = _local_sums_cuda(f, truth.to(fields.device, torch.float32).reshape(C * T, H, W), torch.as_tensor(lat_weights))
    return _tables_from_sums(sums, counts, H * W, C, T, sst_channel)


def plane_shard(n_planes: int, rank: int, world: int) -> range:
    lo, hi = (rank * n_planes) // world, ((rank + 1) * n_planes) // world
    return range(lo, hi)


@torch.no_grad()
def ensemble_metrics_distributed(fields_local: torch.Tensor, truth: torch.Tensor, lat_weights=None, group=None,
                                 sst_channel: Optional[int] = SST_CHANNEL_IDX,
                                 local_sums_fn: Optional[Callable] = None) -> Dict[str, torch.Tensor]:
    """Members are sharded over ranks (`fields_local` [M_r, C, T, H, W] with possibly different M_r per rank); CRPS
    spread and the ensemble mean need all members per grid point, so the fields are re-sharded once — rank r receives
    every member's values for its contiguous slice of the C*T (channel, lead) planes (NCCL all-to-all; an
    all-gather based exchange on backends without all-to-all) — reduced locally, and the [4, planes] partial tables
    are all-gathered.  Every rank returns the full [C, T] tables."""
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
    f = fields_local.to(torch.float32).reshape(M_r, N, H * W)
    counts_m = torch.zeros(world, dtype=torch.int64, device=dev)
    counts_m[rank] = M_r
    dist.all_reduce(counts_m, group=group)
    members = [int(v) for v in counts_m.tolist()]
    mine = plane_shard(N, rank, world)
    n_mine = len(mine)
    backend = dist.get_backend(group)
    if backend == "nccl":
        # send buffer: for destination q, my members' values on q's planes, [M_r, n_q, HW] each, concatenated
        send = torch.cat([f[:, plane_shard(N, q, world).start : plane_shard(N, q, world).stop].reshape(-1) for q in range(world)])
        in_split = [M_r * len(plane_shard(N, q, world)) * H * W for q in range(world)]
        out_split = [members[q] * n_mine * H * W for q in range(world)]
        recv = torch.empty(sum(out_split), dtype=torch.float32, device=dev)
        dist.all_to_all_single(recv, send, out_split, in_split, group=group)
        parts, off = [], 0
        for q in range(world):
            parts.append(recv[off : off + out_split[q]].reshape(members[q], n_mine, H * W))
            off += out_split[q]
        gathered = torch.cat(parts, dim=0)
    else:
        m_max = max(members)
        pad = torch.zeros((m_max, N, H * W), dtype=torch.float32, device=dev)
        pad[:M_r] = f
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        gathered = torch.cat([bufs[q][: members[q], mine.start : mine.stop] for q in range(world)], dim=0)
    t_mine = truth.to(dev, torch.float32).reshape(N, H, W)[mine.start : mine.stop].contiguous()
    sums_l, counts_l = local_sums_fn(gathered.reshape(-1, n_mine, H, W).contiguous(), t_mine, lat_weights)
    n_max = max(len(plane_shard(N, q, world)) for q in range(world))
    packed = torch.zeros((8, n_max), dtype=torch.float64, device=dev)
    packed[:4, :n_mine], packed[4:, :n_mine] = sums_l, counts_l
    allp = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(allp, packed, group=group)
    sums = torch.cat([allp[q][:4, : len(plane_shard(N, q, world))] for q in range(world)], dim=1)
    counts = torch.cat([allp[q][4:, : len(plane_shard(N, q, world))] for q in range(world)], dim=1)
    return _tables_from_sums(sums, counts, H * W, C, T, sst_channel)
