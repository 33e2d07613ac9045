"""
The four staged CUDA kernels of the multi-GPU path (hcu_map2phase / hcu_phase2alm /
hcu_alm2phase / hcu_phase2map with m lists and ring-pair ranges) on ONE GPU: W virtual
ranks are evaluated one after the other and their blocks are moved by hand exactly as the
all-to-all of heracles_b200.dist moves them; parity against the CPU oracle at 1e-10.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("spin,W,nside,lmax", [(0, 1, 16, 32), (0, 2, 32, 64), (2, 3, 32, 80), (0, 4, 64, 128), (2, 2, 128, 256)])
def test_staged_kernels_virtual_ranks(ctx, oracle, spin, W, nside, lmax):
    import torch

    from heracles_b200.dist import ShardPlan, StagedKernels

    plan = ShardPlan(nside, lmax, W)
    k = StagedKernels(ctx, nside, lmax)
    k.sync_streams()
    dev = torch.device("cuda", ctx.device)
    nb = 3 if spin == 0 else 4
    rng = np.random.default_rng(11 + nside + spin)
    full = rng.standard_normal((nb, plan.npix))
    maps = torch.from_numpy(full).to(dev)
    m_all = torch.from_numpy(plan.m_all.copy()).to(dev)
    mpos = torch.from_numpy(plan.mpos.copy()).to(dev)
    mlists = [torch.from_numpy(plan.mlists[g].copy()).to(dev) for g in range(W)]
    per = nb * 4
    # ---- analysis: FFT per ring block, "all-to-all", Legendre per (owner of m, source block)
    send = []
    for g in range(W):
        lo, hi = plan.rp_range(g)
        buf = torch.full(((lmax + 1) * (hi - lo) * per,), np.nan, dtype=torch.float64, device=dev)
        k.map2phase(maps, lo, hi, m_all, buf)
        send.append(buf.view(lmax + 1, hi - lo, nb, 4))
    alms = []
    for g in range(W):
        alm = torch.zeros(nb, plan.nalm, dtype=torch.complex128, device=dev)
        for s in range(W):
            lo, hi = plan.rp_range(s)
            block = send[s][plan.m_off[g]:plan.m_off[g + 1]].contiguous()
            k.phase2alm(block, spin, nb, mlists[g], lo, hi, alm)
        alms.append(alm)
    # the same analysis with all source blocks in ONE launch (hcu_phase2alm_blocks)
    for g in range(W):
        cat = torch.cat([send[s][plan.m_off[g]:plan.m_off[g + 1]].reshape(-1) for s in range(W)])
        alm1 = torch.zeros(nb, plan.nalm, dtype=torch.complex128, device=dev)
        k.phase2alm_blocks(cat, spin, nb, mlists[g], list(plan.rp_bounds), alm1)
        torch.cuda.synchronize()
        d = (alm1 - alms[g]).abs().max().item()
        assert d <= 1e-13 * max(alms[g].abs().max().item(), 1e-300), d
    torch.cuda.synchronize()
    got = sum(a.cpu().numpy() for a in alms)
    ref = oracle.map2alm(nside, lmax, full, spin=spin)
    for c in range(nb):
        assert relerr(got[c], ref[c]) < TOL
    # the m of different owners do not overlap
    nz = [np.abs(a.cpu().numpy()).sum(axis=0) > 0 for a in alms]
    assert not np.any(np.sum(nz, axis=0) > 1)
    # ---- synthesis: Legendre per (owner of m, destination block), "all-to-all", inverse FFT per block
    out = torch.full((nb, plan.npix), np.nan, dtype=torch.float64, device=dev)
    for d in range(W):
        lo, hi = plan.rp_range(d)
        recv = torch.empty(lmax + 1, hi - lo, nb, 4, dtype=torch.float64, device=dev)
        for s in range(W):
            nm = len(plan.mlists[s])
            blk = torch.full((nm, hi - lo, nb, 4), np.nan, dtype=torch.float64, device=dev)
            k.alm2phase(alms[s], spin, nb, mlists[s], lo, hi, blk)
            recv[plan.m_off[s]:plan.m_off[s + 1]] = blk
        k.phase2map(recv, nb, mpos, lo, hi, out)
    # hcu_alm2phase_blocks writes the blocks for all destinations at once
    for s_ in range(W):
        nm = len(plan.mlists[s_])
        allblk = torch.full((nm * plan.nrp * nb * 4,), np.nan, dtype=torch.float64, device=dev)
        k.alm2phase_blocks(alms[s_], spin, nb, mlists[s_], list(plan.rp_bounds), allblk)
        off = 0
        for d in range(W):
            lo, hi = plan.rp_range(d)
            one = torch.empty(nm, hi - lo, nb, 4, dtype=torch.float64, device=dev)
            k.alm2phase(alms[s_], spin, nb, mlists[s_], lo, hi, one)
            n = one.numel()
            assert torch.equal(allblk[off:off + n], one.reshape(-1))
            off += n
    torch.cuda.synchronize()
    back = out.cpu().numpy()
    assert not np.isnan(back).any()  # the blocks cover every pixel
    refmap = oracle.alm2map(nside, lmax, ref, spin=spin)
    assert relerr(back, refmap) < TOL


@pytest.mark.parametrize("spin", [0, 2])
def test_distributed_transform_world1(hb, ctx, oracle, spin):
    import torch

    from heracles_b200.dist import DistributedTransform, ShardPlan, StagedKernels

    nside, lmax, niter = 32, 64, 3
    plan = ShardPlan(nside, lmax, 1)
    k = StagedKernels(ctx, nside, lmax)
    k.sync_streams()
    dev = torch.device("cuda", ctx.device)
    rng = np.random.default_rng(3)
    nb = 14 if spin == 0 else 10  # more than one Legendre batch
    full = rng.standard_normal((nb, plan.npix))
    tr = DistributedTransform(k, plan, 0, niter=niter, device=dev)
    alm = torch.zeros(nb, plan.nalm, dtype=torch.complex128, device=dev)
    fl = 1.0 / (1.0 + 0.01 * np.arange(lmax + 1))
    tr.map2alm(torch.from_numpy(full).to(dev), spin, alm, fl=fl)
    torch.cuda.synchronize()
    ref = oracle.almxfl(oracle.map2alm(nside, lmax, full, spin=spin, niter=niter), fl)
    for c in range(nb):
        assert relerr(alm[c].cpu().numpy(), ref[c]) < TOL


@pytest.mark.parametrize("world", [2, 3])
def test_peer_memory_exchange_between_processes(world):
    """The exchange of the multi-GPU path without a collective: `world` PROCESSES (here sharing one GPU, gloo for the
    barrier) open each other's phase buffers over CUDA IPC; the ring-FFT emission and the Legendre-synthesis flush
    write their rows straight into the consumer's buffer (hcu_map2phase_peers / hcu_alm2phase_peers).  The spectra of
    3 spin-0 maps + 5 spin-2 fields, niter 3, must match the single-process transform (tools/dist_check.py)."""
    import os
    import socket
    import subprocess
    import sys

    from conftest import ROOT

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py"), "--nside", "64", "--niter", "3",
           "--backend", "gloo", "--same-device"]
    env = dict(os.environ, HCU_DIST_EXCHANGE="peer", OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    text = out.stdout + out.stderr
    assert out.returncode == 0 and "-> OK" in text and "through peer memory" in text, text[-3000:]
