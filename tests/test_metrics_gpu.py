"""GPU parity of the on-device metric reduction (lc_metrics_*) against the oracle and the reference-generated golden."""
import os

import numpy as np
import pytest
import torch

from oracle import ladcast_oracle as O

pytestmark = pytest.mark.gpu


def _seeded(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator("cpu").manual_seed(seed)) * scale


def test_metrics_vs_golden(golden_dir):
    from ladcast_b200.evaluate.utils import ensemble_metrics

    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    dec = _seeded((5, 84, 2, 120, 16), 108)
    ref = _seeded((84, 2, 120, 16), 109)
    ref[82, :, 5:9, 3:7] = float("nan")
    tabs = ensemble_metrics(dec.cuda(), ref.cuda())
    for k in ("ens_mse", "crps_skill", "crps_spread", "crps"):
        got = tabs[k].cpu().numpy()
        assert np.allclose(got, g[k], rtol=2e-6, atol=1e-9), k


@pytest.mark.parametrize("M", [1, 2, 20, 50])
def test_metrics_members(M):
    from ladcast_b200.evaluate.utils import ensemble_metrics, get_crps, pointwise_crps_skill, pointwise_crps_spread

    f = _seeded((M, 84, 1, 120, 24), 7 + M)
    t = _seeded((84, 1, 120, 24), 8 + M)
    t[82, 0, :3] = float("nan")
    want = O.ensemble_metrics(f, t)
    got = ensemble_metrics(f.cuda(), t.cuda())
    for k in want:
        assert np.allclose(got[k].cpu().numpy(), want[k].numpy(), rtol=5e-6, atol=1e-9, equal_nan=True), (k, M)
    # pointwise drop-ins (evaluate/utils.py:52-118)
    fc, tc = f[:, :, 0].cuda(), t[:, 0].cuda()
    sp = pointwise_crps_spread(fc, ensemble_dim=0)
    assert torch.allclose(sp.cpu(), O.crps_spread_pointwise(f[:, :, 0]), rtol=1e-5, atol=1e-6)
    sk = pointwise_crps_skill(fc, tc.unsqueeze(0), 0)
    assert torch.allclose(sk.cpu(), torch.abs(t[:, 0].unsqueeze(0) - f[:, :, 0]).mean(0), rtol=1e-5, atol=1e-6, equal_nan=True)
    cr = get_crps(fc, tc.unsqueeze(0), 0)
    assert torch.allclose(cr.cpu(), sk.cpu() - 0.5 * sp.cpu(), rtol=1e-5, atol=1e-6, equal_nan=True)


def test_metrics_full_size_properties():
    """BASELINE-size planes (120x240), ens=20: spread is translation invariant and scales linearly; skill of a
    forecast equal to the truth is 0; identical members give zero spread and crps == skill == |error|."""
    from ladcast_b200.evaluate.utils import ensemble_metrics

    f = _seeded((20, 84, 2, 120, 240), 11).cuda()
    t = _seeded((84, 2, 120, 240), 12).cuda()
    a = ensemble_metrics(f, t)
    b = ensemble_metrics(f * 3.0 + 5.0, t * 3.0 + 5.0)
    assert torch.allclose(b["crps_spread"], 3.0 * a["crps_spread"], rtol=1e-5)
    assert torch.allclose(b["crps"], 3.0 * a["crps"], rtol=1e-4)
    assert torch.allclose(b["ens_mse"], 9.0 * a["ens_mse"], rtol=1e-4)
    same = t.unsqueeze(0).expand(20, -1, -1, -1, -1).contiguous()
    z = ensemble_metrics(same, t)
    # the sorted rank-weighted sum (the reference's formulation, evaluate/utils.py:86-99) cancels to rounding, not to 0
    assert float(z["crps_spread"].abs().max()) < 1e-5 and float(z["crps_skill"].abs().max()) == 0.0


def test_get_acc_vs_oracle():
    """get_acc (evaluate/utils.py:122-149): lat-weighted and unweighted, with NaNs in truth (nanmean semantics)."""
    from ladcast_b200.evaluate.utils import get_acc

    f = _seeded((84, 120, 24), 21)
    t = _seeded((84, 120, 24), 22)
    c = _seeded((84, 120, 24), 23, 0.3)
    t[82, :4] = float("nan")
    w = torch.from_numpy(O.lat_weights(120)).view(-1, 1)
    want_w, want_u = O.get_acc(f, t, c, w), O.get_acc(f, t, c)  # pinned to the reference function by tests/golden/acc.npz
    got_w = get_acc(f.cuda(), t.cuda(), c.cuda(), w.cuda())
    got_u = get_acc(f.cuda(), t.cuda(), c.cuda())
    assert torch.allclose(got_w.cpu(), want_w, rtol=1e-6, atol=1e-9)
    assert torch.allclose(got_u.cpu(), want_u, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("M", [1, 7, 20, 50])
def test_metrics_pointer_form_matches_strided(M):
    """lc_metrics_accumulate_ptrs (one base pointer per member: the form that reads peer-GPU members in place) against
    lc_metrics_accumulate_strided on the same values, with the members scattered over separate allocations and a plane
    offset as in the distributed reduction: bit-identical tables."""
    import ctypes

    from ladcast_b200 import _lib
    from ladcast_b200.evaluate.utils import get_normalized_lat_weights_based_on_cos

    lib = _lib.load()
    N, H, W, n0, n_mine = 12, 120, 48, 3, 5  # a rank's slice [n0, n0 + n_mine) of N planes
    fields = _seeded((M, N, H, W), 300 + M).cuda()
    truth = _seeded((N, H, W), 301).cuda()
    truth[4, 7:9, 5:11] = float("nan")
    lw = torch.from_numpy(get_normalized_lat_weights_based_on_cos(np.linspace(-88.5, 90, H))).cuda()
    blocks = [fields[m].clone() for m in range(M)]  # every member in its own allocation
    ptrs = (ctypes.c_void_p * M)(*[b.data_ptr() + 4 * n0 * H * W for b in blocks])
    t_mine = truth[n0 : n0 + n_mine].contiguous()
    s1, c1 = [torch.empty((4, n_mine), device="cuda", dtype=torch.float64) for _ in range(2)]
    s2, c2 = [torch.empty((4, n_mine), device="cuda", dtype=torch.float64) for _ in range(2)]
    _lib.check(lib.lc_metrics_accumulate_ptrs(ptrs, _lib.ptr(t_mine), _lib.ptr(lw), M, n_mine, H, W, _lib.ptr(s1),
                                              _lib.ptr(c1), _lib.stream()), "lc_metrics_accumulate_ptrs")
    sl = fields[:, n0 : n0 + n_mine]
    _lib.check(lib.lc_metrics_accumulate_strided(_lib.ptr_any(sl), fields.stride(0), _lib.ptr(t_mine), _lib.ptr(lw), M,
                                                 n_mine, H, W, _lib.ptr(s2), _lib.ptr(c2), _lib.stream()),
               "lc_metrics_accumulate_strided")
    torch.cuda.synchronize()
    assert torch.equal(c1, c2)
    assert torch.allclose(s1, s2, rtol=1e-12, atol=0, equal_nan=True)  # fp64 atomics: order of block sums may differ
    with pytest.raises(_lib.LadcastB200Error):
        _lib.check(lib.lc_metrics_accumulate_ptrs(ptrs, _lib.ptr(t_mine), _lib.ptr(lw), 65, n_mine, H, W, _lib.ptr(s1),
                                                  _lib.ptr(c1), _lib.stream()), "lc_metrics_accumulate_ptrs")


def test_peer_memory_exchange_world1():
    """exchange="p2p" end to end in a one-rank NCCL group: PeerBuffer (lc_ipc_alloc, torch view of the raw pointer),
    fields produced directly in the peer buffer (no staging copy) and elsewhere (one local copy), pointer-form kernel —
    all equal to ensemble_metrics; the multi-rank case is tools/dist_metrics_check.py (profiles/r02_dist_metrics_2gpu.jsonl)."""
    import torch.distributed as dist

    from ladcast_b200.evaluate import utils as U

    if dist.is_initialized():
        pytest.skip("a process group already exists")
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{29650 + os.getpid() % 300}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        M, C, T, H, W = 6, 84, 2, 120, 16
        fields = _seeded((M, C, T, H, W), 410).cuda()
        truth = _seeded((C, T, H, W), 411).cuda()
        truth[82, :, 3:6, 2:5] = float("nan")
        want = U.ensemble_metrics(fields, truth)
        tm = {}
        got = U.ensemble_metrics_distributed(fields, truth, exchange="p2p", timings=tm)  # copied into the peer buffer
        buf = U._peer_buffer(M * C * T * H * W, fields.device)
        inplace = buf.tensor[: fields.numel()].view(M, C, T, H, W)
        inplace.copy_(fields * 2.0)
        got2 = U.ensemble_metrics_distributed(inplace, truth, exchange="p2p")  # already there: read in place
        want2 = U.ensemble_metrics(fields * 2.0, truth)
        for k in want:
            assert torch.allclose(got[k], want[k], rtol=1e-12, atol=0, equal_nan=True), k
            assert torch.allclose(got2[k], want2[k], rtol=1e-12, atol=0, equal_nan=True), k
        assert tm["kernel_ms"] > 0
    finally:
        U.release_peer_buffers()
        dist.destroy_process_group()
