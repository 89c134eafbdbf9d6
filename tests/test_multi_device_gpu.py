"""One process driving several GPUs (cudaSetDevice(g) then GPU_NTT_Inplace on that device's batch slice):
per-device state of the engine (shared-memory opt-in of the tuned kernels, cached workspaces) must not leak between
devices.  Skipped on single-GPU boxes."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from gpu_ntt_b200 import capi  # noqa: E402
from gpu_ntt_b200.multigpu import batch_slice  # noqa: E402
from oracle import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits,logn", [(64, 16), (32, 14), (64, 10)])
def test_batch_slices_on_two_devices_in_one_process(bits, logn):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    P = O.merge_params(logn, O.X_N_minus, bits)
    batch = 6
    x = O.example_input(P.modulus, batch << logn, seed=3).reshape(batch, -1)
    want = O.merge_ntt(x, P)
    outs = []
    for g in range(2):
        lo, hi = batch_slice(g, 2, batch)
        with torch.cuda.device(g):
            if bits == 64:
                d = torch.from_numpy(np.ascontiguousarray(x[lo:hi]).view(np.int64)).cuda()
                tab = torch.from_numpy(P.fwd_br.view(np.int64)).cuda()
            else:
                d = torch.from_numpy(np.ascontiguousarray(x[lo:hi]).astype(np.uint32).view(np.int32)).cuda()
                tab = torch.from_numpy(P.fwd_br.astype(np.uint32).view(np.int32)).cuda()
            capi.ntt(d, tab, P.modulus, logn, O.X_N_minus)
            torch.cuda.synchronize()
            a = d.cpu().numpy()
            outs.append(a.view(np.uint64) if bits == 64 else a.view(np.uint32).astype(np.uint64))
    assert (np.concatenate(outs) == want).all()


def test_scatter_transform_gather_over_peer_copies():
    """The whole batch lives on GPU 0: gpuntt_b200_scatter_batch hands every device its slice (cudaMemcpyPeerAsync,
    NVLink where peer access exists), each device transforms its slice, gpuntt_b200_gather_batch brings the results back.
    RNS-aligned slices (mod_count = 3) so that every device could use the same modulus array."""
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs two GPUs")
    from gpu_ntt_b200.multigpu import gather_batch, scatter_batch
    ndev = min(ndev, 4)
    logn, batch, mc = 13, 21, 3
    P = O.merge_params(logn, O.X_N_minus, 64)
    x = O.example_input(P.modulus, batch << logn, seed=8).reshape(batch, -1)
    want = O.merge_ntt(x, P)
    with torch.cuda.device(0):
        src = torch.from_numpy(x.view(np.int64)).cuda()
        back = torch.zeros_like(src)
    parts, tabs = [], []
    for g in range(ndev):
        lo, hi = batch_slice(g, ndev, batch, mc)
        with torch.cuda.device(g):
            parts.append(torch.zeros((hi - lo, 1 << logn), dtype=torch.int64, device=f"cuda:{g}"))
            tabs.append(torch.from_numpy(P.fwd_br.view(np.int64)).cuda())
    for g in range(ndev):
        torch.cuda.synchronize(g)
    scatter_batch(src, parts, list(range(ndev)), mod_count=mc)
    for g in range(ndev):
        torch.cuda.synchronize(g)
    for g in range(ndev):
        with torch.cuda.device(g):
            if parts[g].shape[0]:
                capi.ntt(parts[g], tabs[g], P.modulus, logn, O.X_N_minus)
            torch.cuda.synchronize()
    gather_batch(back, parts, list(range(ndev)), mod_count=mc)
    for g in range(ndev):
        torch.cuda.synchronize(g)
    assert (back.cpu().numpy().view(np.uint64) == want).all()
