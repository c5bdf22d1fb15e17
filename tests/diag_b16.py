"""Diagnostic (GPU): where does the batch-16 training step leave the oracle? Per-key loss deviation of the first steps,
per-parameter gradient cosine of G and D at batch 16, per-tensor weight movement after one step."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MMH_VGG19_RANDOM", "1")
import torch  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import patn_ref as O  # noqa: E402
from mmhand_b200.options import make_opt  # noqa: E402

DEV = "cuda"
B = int(os.environ.get("DIAG_B", "16"))
S = 256


def _inputs(B, S, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    d = dict(H1=r(B, 3, S, S) * 2 - 1, P1=(r(B, 21, S, S) > 0.98).float() * r(B, 21, S, S), D1=r(B, 3, S, S) * 2 - 1,
             H2=r(B, 3, S, S) * 2 - 1, P2=(r(B, 21, S, S) > 0.98).float() * r(B, 21, S, S), D2=r(B, 3, S, S) * 2 - 1)
    return {k: v.to(DEV) for k, v in d.items()}


def _sd(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def grads():
    from models.Discriminator import Discriminator
    from models.Generator import Generator
    from models.network_utils import get_norm_layer, init_weights
    torch.manual_seed(49)
    g = Generator([3, 42, 6], 3, 64, get_norm_layer('batch'), True, 9)
    init_weights(g, 'normal')
    g = g.to(DEV)
    b = _inputs(B, S, 2)
    x = [b["H1"], torch.cat((b["P1"], b["P2"]), 1), torch.cat((b["D1"], b["D2"]), 1)]
    sd = _sd(g)
    g.train()
    g._step = 0
    y = g(x)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).to(DEV)
    y.backward(gy)
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
           for k, v in sd.items()}
    want = O.generator_forward(sdo, x, train=True, use_dropout=True, drop=O.DropCtx("hash", 0, 0, 0))
    want.backward(gy)
    print("G B=%d fwd max-abs %.4f" % (B, (y.detach() - want.detach()).abs().max().item()))
    cos = sorted((torch.nn.functional.cosine_similarity(p.grad.flatten(), sdo[k].grad.flatten(), dim=0).item(),
                  (p.grad.norm() / sdo[k].grad.norm()).item(), k) for k, p in g.named_parameters())
    print("G grad (cos, norm ratio) worst 12:")
    for c in cos[:12]:
        print("   %.4f %.4f %s" % c)
    print("   median", cos[len(cos) // 2])
    del g, want, sdo
    torch.cuda.empty_cache()
    d = Discriminator(24, 64, get_norm_layer('batch'), True, 3, [], 'reflect', False, 2)
    init_weights(d, 'normal')
    d = d.to(DEV)
    xd = (torch.rand(B, 24, S, S, generator=torch.Generator().manual_seed(5)) * 2 - 1).to(DEV).requires_grad_(True)
    sd = _sd(d)
    d.train()
    d._step, d.drop_net_id = 0, 2
    yd = d(xd)
    from models.network_utils import GANLoss
    GANLoss(use_lsgan=False)(yd, False).backward()
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
           for k, v in sd.items()}
    xo = xd.detach().clone().requires_grad_(True)
    yo = O.discriminator_forward(sdo, xo, True, True, drop=O.DropCtx("hash", 0, 0, 2))
    O.gan_loss(yo, False).backward()
    print("D B=%d fwd max-abs %.4f input-grad cos %.4f" % (
        B, (yd.detach() - yo.detach()).abs().max().item(),
        torch.nn.functional.cosine_similarity(xd.grad.flatten(), xo.grad.flatten(), dim=0).item()))
    cos = sorted((torch.nn.functional.cosine_similarity(p.grad.flatten(), sdo[k].grad.flatten(), dim=0).item(),
                  (p.grad.norm() / sdo[k].grad.norm()).item(), k) for k, p in d.named_parameters())
    print("D grad (cos, norm ratio) worst 8:")
    for c in cos[:8]:
        print("   %.4f %.4f %s" % c)


def steps(n=3, use_tape=True):
    from models.MMHandModel import MMHandModel
    torch.manual_seed(49)
    random.seed(49)
    opt = make_opt(batchSize=B, fineSize=S, pool_size=50, local_rank=0, gpu=0, seed=49)
    m = MMHandModel(opt)
    m.master = False
    m.use_tape = use_tape
    vsd = {k: v.detach().clone() for k, v in m.criterionL1.vgg_submodel.state_dict().items()}
    tr = O.OracleTrainer(_sd(m.netG), _sd(m.netD_PB), _sd(m.netD_PP), vsd, opt.lambda_A, opt.lambda_B,
                         opt.lambda_GAN, opt.lr, opt.beta1, opt.pool_size, True, True, dropout="hash", seed=49,
                         device=DEV)
    for i in range(n):
        b = _inputs(B, S, 100 + i)
        st = random.getstate()
        m.set_input(b)
        m.optimize_parameters()
        mine = {k: float(v) for k, v in m.get_current_errors().items()}
        random.setstate(st)
        ref = tr.step(b["H1"], b["P1"], b["D1"], b["H2"], b["P2"], b["D2"])
        print("step %d (tape=%s):" % (i, use_tape), {k: "%.4f/%.4f" % (mine[k], ref[k]) for k in ref})
        for name, net, sd in (("G", m.netG, tr.g), ("D_PP", m.netD_PP, tr.dpp), ("D_PB", m.netD_PB, tr.dpb)):
            worst = sorted(((p.detach() - sd[k].detach()).abs().mean().item() / opt.lr, k)
                           for k, p in net.named_parameters())[-3:]
            print("   %s weight mean|diff|/lr worst:" % name, [("%.3f" % a, k) for a, k in worst])


if __name__ == "__main__":
    what = sys.argv[1:] or ["grads", "steps", "steps_eager"]
    if "grads" in what:
        grads()
    if "steps" in what:
        steps(3, True)
    if "steps_eager" in what:
        steps(3, False)
