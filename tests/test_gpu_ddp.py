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


def _joint_worker(rank, world, port, mode, out_dir, S, ngf, steps):
    """One rank of the joint-batch test: per-rank batch 1 = sample `rank` of the joint batch, dropout on."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MMH_SYNCBN"] = mode
    os.environ["MMH_VGG19_RANDOM"] = "1"
    os.environ.setdefault("WORLD_SIZE", str(world))
    torch.cuda.set_device(rank)
    import torch.distributed as dist
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from models.MMHandModel import MMHandModel
    from mmhand_b200.options import make_opt
    torch.manual_seed(5 + 100 * rank)          # rank-dependent init: rank 0's is broadcast (apex DDP semantics)
    random.seed(5)
    opt = make_opt(batchSize=1, fineSize=S, ngf=ngf, ndf=ngf, pool_size=0, local_rank=rank, gpu=rank, seed=7,
                   distributed=True)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = MMHandModel(opt)
    m.master = False
    sd = lambda net: {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    res = {"peer": m.world.peer is not None, "init": {"g": sd(m.netG), "dpb": sd(m.netD_PB), "dpp": sd(m.netD_PP),
                                                      "vgg": sd(m.criterionL1.vgg_submodel)}}
    g = torch.Generator().manual_seed(321)
    r = lambda *s: torch.rand(*s, generator=g)
    errs = []
    for _ in range(steps):
        b = dict(H1=r(world, 3, S, S) * 2 - 1, P1=r(world, 21, S, S), D1=r(world, 3, S, S) * 2 - 1,
                 H2=r(world, 3, S, S) * 2 - 1, P2=r(world, 21, S, S), D2=r(world, 3, S, S) * 2 - 1)
        m.set_input({k: v[rank:rank + 1] for k, v in b.items()})
        m.optimize_parameters()
        errs.append({k: float(v) for k, v in m.get_current_errors().items()})
    res["errs"] = errs
    res["g"] = sd(m.netG)
    torch.save(res, os.path.join(out_dir, "joint_%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def _joint_vs_oracle(world, mode, S=64, ngf=32, steps=3):
    """N ranks (one sample each, SyncBN + gradient all-reduce, joint-batch dropout masks) against the ORACLE stepping
    on the joint batch of N samples on one GPU: the data-parallel path computes the single-process result."""
    from oracle import patn_ref as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with tempfile.TemporaryDirectory() as td:
        port = 29700 + (os.getpid() % 250) + 7 * world
        mp.spawn(_joint_worker, args=(world, port, mode, td, S, ngf, steps), nprocs=world, join=True)
        res = [torch.load(os.path.join(td, "joint_%d.pt" % r)) for r in range(world)]
    assert all(r["peer"] == (mode == "peer") for r in res)
    init = res[0]["init"]
    for r in res[1:]:           # rank 0's initial state reached every rank although they were seeded differently
        for part in ("g", "dpb", "dpp", "vgg"):
            assert all(torch.equal(init[part][k], r["init"][part][k]) for k in init[part]), part
    for r in res[1:]:           # replicas in lock-step after the steps
        assert all(torch.equal(res[0]["g"][k], r["g"][k]) for k in res[0]["g"])
    tr = O.OracleTrainer(init["g"], init["dpb"], init["dpp"], init["vgg"], 10.0, 10.0, 5.0, 2e-4, 0.5, 0, True, True,
                         dropout="hash", seed=7, device="cuda")
    g = torch.Generator().manual_seed(321)
    r_ = lambda *s: torch.rand(*s, generator=g)
    worst = 0.0
    for i in range(steps):
        b = dict(H1=r_(world, 3, S, S) * 2 - 1, P1=r_(world, 21, S, S), D1=r_(world, 3, S, S) * 2 - 1,
                 H2=r_(world, 3, S, S) * 2 - 1, P2=r_(world, 21, S, S), D2=r_(world, 3, S, S) * 2 - 1)
        b = {k: v.cuda() for k, v in b.items()}
        ref = tr.step(b["H1"], b["P1"], b["D1"], b["H2"], b["P2"], b["D2"])
        for k in ref:
            mine = sum(r["errs"][i][k] for r in res) / world        # mean over the joint batch = mean of rank means
            rel = abs(mine - ref[k]) / max(abs(ref[k]), 1e-6)
            worst = max(worst, rel)
            assert rel <= 1e-2, (world, mode, i, k, mine, ref[k])
    print("joint-batch vs oracle, world %d (%s): worst relative loss deviation %.3g" % (world, mode, worst))
    return worst


@pytest.mark.timeout(900)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["nccl", "peer"])
def test_two_ranks_match_oracle_on_joint_batch(mode):
    _joint_vs_oracle(2, mode)


@pytest.mark.timeout(900)
@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs four GPUs")
@pytest.mark.parametrize("mode", ["nccl", "peer"])
def test_all_ranks_match_oracle_on_joint_batch(mode):
    """4 GPUs, or 8 when the box has them."""
    _joint_vs_oracle(8 if torch.cuda.device_count() >= 8 else 4, mode)
