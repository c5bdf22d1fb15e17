"""Data-parallel path on two GPUs (NCCL + the peer-memory SyncBN exchange of csrc/peer.cu).

(1) mmh_peer_sum against NCCL all-reduce over hundreds of back-to-back exchanges (slot reuse, sequence numbers);
(2) three MMHandModel optimisation steps with the exchange fused into the BN finalise kernels against the same steps
    with one NCCL all-reduce per exchange (MMH_SYNCBN=nccl): with two ranks a + b is the same in either order, so
    losses and weights must agree to accumulation noise of the atomics only;
(3) both ranks end with identical weights (replicas stay in lock-step).
Skipped on boxes with fewer than two GPUs; the host-side logic of the N > 1 path is covered on CPU by
tests/test_ddp_gloo.py."""
import os
import random
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MMH_SYNCBN"] = mode
    os.environ.setdefault("WORLD_SIZE", str(world))      # as under torchrun (multi-process defaults of the library)
    torch.cuda.set_device(rank)
    import torch.distributed as dist
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from mmhand_b200 import runtime
    from models.MMHandModel import MMHandModel
    from mmhand_b200.options import make_opt
    torch.manual_seed(5)
    random.seed(5)
    opt = make_opt(batchSize=1, fineSize=64, ngf=32, ndf=32, pool_size=0, local_rank=rank, gpu=rank, seed=7,
                   distributed=True)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = MMHandModel(opt)
    m.master = False
    res = {"peer": m.world.peer is not None}
    if mode == "peer":
        assert m.world.peer is not None, "peer mailboxes were not connected"
        ops = runtime.get_ops(torch.device("cuda", rank))
        g = torch.Generator().manual_seed(100 + rank)
        worst = 0.0
        for i in range(300):
            n = (1, 7, 128, 1024, 2048)[i % 5]
            x = torch.randn(n, generator=g).cuda()
            want = x.clone()
            dist.all_reduce(want)
            ops.peer_sum(m.world, x)
            worst = max(worst, (x - want).abs().max().item())
        res["peer_sum_err"] = worst
    g = torch.Generator().manual_seed(123)
    r = lambda *s: torch.rand(*s, generator=g)
    errs = []
    for _ in range(3):
        b = dict(H1=r(2, 3, 64, 64) * 2 - 1, P1=r(2, 21, 64, 64), D1=r(2, 3, 64, 64) * 2 - 1,
                 H2=r(2, 3, 64, 64) * 2 - 1, P2=r(2, 21, 64, 64), D2=r(2, 3, 64, 64) * 2 - 1)
        m.set_input({k: v[rank:rank + 1] for k, v in b.items()})
        m.optimize_parameters()
        errs.append({k: float(v) for k, v in m.get_current_errors().items()})
    res["errs"] = errs
    res["g"] = {k: v.detach().cpu().clone() for k, v in m.netG.state_dict().items()}
    res["d"] = {k: v.detach().cpu().clone() for k, v in m.netD_PB.state_dict().items()}
    torch.save(res, os.path.join(out_dir, "%s_%d.pt" % (mode, rank)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_syncbn_matches_nccl_on_two_gpus():
    with tempfile.TemporaryDirectory() as td:
        port = 29600 + (os.getpid() % 300)
        for i, mode in enumerate(("peer", "nccl")):
            mp.spawn(_worker, args=(2, port + i, mode, td), nprocs=2, join=True)
        peer = [torch.load(os.path.join(td, "peer_%d.pt" % r)) for r in range(2)]
        nccl = [torch.load(os.path.join(td, "nccl_%d.pt" % r)) for r in range(2)]
    assert peer[0]["peer"] and peer[1]["peer"] and not nccl[0]["peer"]
    assert max(p["peer_sum_err"] for p in peer) == 0.0           # two ranks: a + b in either order
    for part in ("g", "d"):
        for k in peer[0][part]:
            # replicas in lock-step: rank 0 and rank 1 hold the same weights after three steps
            assert torch.equal(peer[0][part][k], peer[1][part][k]), (part, k)
    lr = 2e-4
    for part in ("g", "d"):
        for k, a in peer[0][part].items():
            b = nccl[0][part][k]
            if not a.is_floating_point():
                assert torch.equal(a, b), k
                continue
            d = (a - b).abs()
            if k.endswith("running_mean") or k.endswith("running_var"):
                # the exchanged statistics themselves (after three steps of slightly diverging weights and bf16
                # activations; the two modes also sum in different orders: conv-epilogue vs statistics kernel)
                assert d.max().item() <= 3e-2 * a.abs().max().item() + 1e-4, (part, k, d.max().item())
                continue
            # same arithmetic up to the order of fp32 atomics inside the reduction kernels (Adam normalises the
            # update: a weight whose gradient is ~0 may move by up to ~2 lr per step in either run; measured: a
            # scale-invariant stem weight differs by 0.36 lr on average after three steps between two such runs)
            assert d.max().item() <= 2 * 3 * lr + 1e-3 * a.abs().max().item(), (part, k, d.max().item())
            assert d.mean().item() <= 1.5 * lr + 1e-3 * a.abs().mean().item() + 1e-6, (part, k, d.mean().item())
    for ea, eb in zip(peer[0]["errs"], nccl[0]["errs"]):
        for k in ea:
            assert abs(ea[k] - eb[k]) <= 2e-2 * max(abs(eb[k]), 1e-3), (k, ea[k], eb[k])
