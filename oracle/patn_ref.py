"""ORACLE (test infrastructure, never imported by the product path).

Functional fp32 restatement, in plain torch ops, of the arithmetic of the MM-HAND hot path, keyed by the
reference's own ``state_dict`` names so that the same weights drive the reference modules, this oracle and
the CUDA path:

  generator_forward      models/Generator.py:115-130 (PATBlock.forward, incl. the swapped return order that the
                         caller unpacks as x1, x2, x3 -> pose/depth streams alternate, Generator.py:130 vs :278),
                         :269-283 (PATNModel.forward), layer stacks :158-259
  generator2_forward     baselines/quantitative_on_benchmarks/networks/model_variants.py:8-173 (two-stream PATNetwork)
  ssim                   baselines/quantitative_on_benchmarks/pytorch_ssim/__init__.py:7-73
  discriminator_forward  models/Discriminator.py:53-55 (ResnetBlock.forward), :79-154
  gan_loss               models/network_utils.py:129-163 (always BCEWithLogitsLoss, :141)
  l1_plus_perceptual     losses/L1_plus_perceptualLoss.py:32-75
  OracleTrainer          models/MMHandModel.py:215-221 (forward), :236-261 (backward_G), :263-292 (backward_D_*),
                         :310-330 (optimize_parameters: G, then D_PP, then D_PB), Adam at :90-98
  ImagePoolRef           util/image_pool.py:14-34

Pinned against the reference modules themselves by the golden vectors in tests/golden/ generated from the live
reference by oracle/make_golden.py (tests/test_oracle_golden.py).

Dropout: the reference draws nn.Dropout masks from torch's global generator (Generator.py:76-77,
Discriminator.py:33-34); for a reproducible three-way comparison the oracle takes the masks from
``dropout_mask`` below, the same counter-based hash the CUDA kernels evaluate.
"""
import random

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOM = 0.1


# ------------------------------------------------------------------------------------------------ dropout
def _mix32(x):
    """murmur3 finaliser on uint32 carried in int64 tensors."""
    m = 0xFFFFFFFF
    x = x & m
    x = x ^ (x >> 16)
    x = (x * 0x85EBCA6B) & m
    x = x ^ (x >> 13)
    x = (x * 0xC2B2AE35) & m
    x = x ^ (x >> 16)
    return x


def dropout_key(seed, layer_id, step):
    m = 0xFFFFFFFF
    k = (seed * 0x9E3779B1 + layer_id * 0x85EBCA77 + step * 0xC2B2AE3D + 0x27D4EB2F) & m
    return k


def dropout_mask(shape, key, device="cpu"):
    """keep-mask (float 0/1) for an NCHW tensor [B, C, H, W]: one hash word per (pixel, group of 8 channels),
    word = ((b*H + h)*W + w) * ceil(C/8) + c//8; element (b, c, h, w) keeps iff bit (c % 8) of
    mix32(word * 0x9E3779B1 + key) is set (the same function the CUDA kernels evaluate, ew_framework.h::drop_bits)."""
    B, C, H, W = shape
    G = (C + 7) // 8
    pix = torch.arange(B * H * W, dtype=torch.int64, device=device).view(B, 1, H, W)
    c = torch.arange(C, dtype=torch.int64, device=device).view(1, C, 1, 1)
    word = pix * G + c // 8
    h = _mix32(word * 0x9E3779B1 + key)
    return ((h >> (c % 8)) & 1).to(torch.float32)


class DropCtx:
    """Hands out per-layer dropout masks. mode: 'off' | 'hash'."""

    def __init__(self, mode="off", seed=0, step=0, net_id=0):
        self.mode, self.seed, self.step, self.net_id = mode, seed, step, net_id
        self.counter = 0

    def apply(self, x):
        lid = self.net_id * 1000 + self.counter
        self.counter += 1
        if self.mode == "off":
            return x
        m = dropout_mask(tuple(x.shape), dropout_key(self.seed, lid, self.step), x.device)
        return x * m * 2.0


# ------------------------------------------------------------------------------------------------ layers
def _bn(sd, prefix, x, train):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=train, momentum=BN_MOM, eps=BN_EPS)


def _rpad(x, p):
    return F.pad(x, (p, p, p, p), mode="reflect")


# Optional storage-rounding hook (tests only): when set to e.g. ``lambda t: t.bfloat16().float()`` the oracle rounds
# conv operands (activations, weights) and raw conv outputs exactly where the CUDA path stores bf16,
# keeping fp32 accumulation, statistics and residual trunks. None = the reference's plain fp32 arithmetic.
QUANT = None
# Selective rounding points for the ablation of tests/diag_quant_ablation.py: any subset of {"act", "w", "raw"}
# (conv input activations / weights / raw conv outputs). None = all three.
QUANT_POINTS = None


def _q(t, point=None):
    if QUANT is None or (QUANT_POINTS is not None and point not in QUANT_POINTS):
        return t
    return QUANT(t)


def _conv(x, w, bias=None, stride=1, padding=0, q_out=True):
    y = F.conv2d(_q(x, "act"), _q(w, "w"), bias, stride=stride, padding=padding)
    return _q(y, "raw") if q_out else y


def _convT(x, w):
    return _q(F.conv_transpose2d(_q(x, "act"), _q(w, "w"), stride=2, padding=1, output_padding=1), "raw")


def _down_stream(sd, p, x, train, n_down=2):
    """pad3-conv7-BN-ReLU, n_down x [conv3 s2 p1 - BN - ReLU]  (Generator.py:158-223, Discriminator.py:79-133: the
    channel counts -- incl. the 4 ndf -> 4 ndf third stage of n_downsampling == 3 -- are those of the weights)."""
    x = F.relu(_bn(sd, p + ".2", _conv(_rpad(x, 3), sd[p + ".1.weight"]), train))
    for i in range(n_down):
        x = F.relu(_bn(sd, p + ".%d" % (5 + 3 * i), _conv(x, sd[p + ".%d.weight" % (4 + 3 * i)], stride=2, padding=1),
                       train))
    return x


def _conv_block(sd, p, x, train, use_dropout, drop, final_bn):
    """pad,conv,BN,ReLU,[drop],pad,conv,[BN]  (Generator.py:40-113, Discriminator.py:14-51)."""
    j = 6 if use_dropout else 5
    h = F.relu(_bn(sd, p + ".2", _conv(_rpad(x, 1), sd[p + ".1.weight"]), train))
    if use_dropout:
        h = drop.apply(h) if train else h
    h = _conv(_rpad(h, 1), sd[p + ".%d.weight" % j])
    if final_bn:
        h = _bn(sd, p + ".%d" % (j + 1), h, train)
    return h


def generator_forward(sd, inputs, train=False, use_dropout=True, n_blocks=9, drop=None, taps=None):
    drop = drop or DropCtx("off")
    x1, x2, x3 = inputs
    x1 = _down_stream(sd, "model.stream1_down", x1, train)
    x2 = _down_stream(sd, "model.stream2_down", x2, train)
    x3 = _down_stream(sd, "model.stream3_down", x3, train)
    if taps is not None:
        taps["down"] = (x1, x2, x3)
    for i in range(n_blocks):
        p = "model.att.%d" % i
        # dropout call order inside PATBlock.forward: stream1, stream2, stream3 (Generator.py:116-118)
        c1 = _conv_block(sd, p + ".conv_block_stream1", x1, train, use_dropout, drop, True)
        o2 = _conv_block(sd, p + ".conv_block_stream2", x2, train, use_dropout, drop, False)
        o3 = _conv_block(sd, p + ".conv_block_stream3", x3, train, use_dropout, drop, False)
        out = x1 + c1 * torch.sigmoid(o2) * torch.sigmoid(o3)
        # returned as (out, cat(x3_out,out), cat(x2_out,out), _) and unpacked as x1, x2, x3 -> swap
        x1, x2, x3 = out, torch.cat((o3, out), 1), torch.cat((o2, out), 1)
        if taps is not None:
            taps["att%d" % i] = out
    u = "model.stream1_up"
    h = _convT(x1, sd[u + ".0.weight"])
    h = F.relu(_bn(sd, u + ".1", h, train))
    h = _convT(h, sd[u + ".3.weight"])
    h = F.relu(_bn(sd, u + ".4", h, train))
    h = _conv(_rpad(h, 3), sd[u + ".7.weight"], sd[u + ".7.bias"], q_out=False)
    return torch.tanh(h)


def generator2_forward(sd, inputs, train=False, use_dropout=True, n_blocks=9, drop=None):
    """Two-stream pose-transfer generator of the benchmark harness
    (baselines/quantitative_on_benchmarks/networks/model_variants.py:58-68 PATBlock.forward, :139-153 PATNModel.forward):
    ``out = x1 + stream1(x1) * sigmoid(stream2(x2))``, next pose input ``cat(x2_out, out)`` -- no depth stream, no swap."""
    drop = drop or DropCtx("off")
    x1, x2 = inputs
    x1 = _down_stream(sd, "model.stream1_down", x1, train)
    x2 = _down_stream(sd, "model.stream2_down", x2, train)
    for i in range(n_blocks):
        p = "model.att.%d" % i
        c1 = _conv_block(sd, p + ".conv_block_stream1", x1, train, use_dropout, drop, True)
        o2 = _conv_block(sd, p + ".conv_block_stream2", x2, train, use_dropout, drop, False)
        out = x1 + c1 * torch.sigmoid(o2)
        x1, x2 = out, torch.cat((o2, out), 1)
    u = "model.stream1_up"
    h = _convT(x1, sd[u + ".0.weight"])
    h = F.relu(_bn(sd, u + ".1", h, train))
    h = _convT(h, sd[u + ".3.weight"])
    h = F.relu(_bn(sd, u + ".4", h, train))
    h = _conv(_rpad(h, 3), sd[u + ".7.weight"], sd[u + ".7.bias"], q_out=False)
    return torch.tanh(h)


def ssim(img1, img2, window_size=11, size_average=True):
    """baselines/quantitative_on_benchmarks/pytorch_ssim/__init__.py:7-39,65-73: Gaussian window (sigma 1.5) built as
    the fp32 outer product of the normalised 1-D window, five zero-padded depthwise convolutions, C1 = 0.01^2,
    C2 = 0.03^2, mean over everything (or per image)."""
    from math import exp
    ch = img1.shape[1]
    g = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    w = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(ch, 1, window_size, window_size).contiguous().to(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, w, padding=pad, groups=ch)
    mu2 = F.conv2d(img2, w, padding=pad, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=pad, groups=ch) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=pad, groups=ch) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=pad, groups=ch) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)


def discriminator_forward(sd, x, train=True, use_dropout=True, n_blocks=3, drop=None, n_downsampling=2):
    """Discriminator.py:79-151: stem, n_downsampling stride-2 stages, n_blocks ResnetBlocks (no head)."""
    drop = drop or DropCtx("off")
    x = _down_stream(sd, "model", x, train, n_downsampling)
    for i in range(n_blocks):
        p = "model.%d.conv_block" % (4 + 3 * n_downsampling + i)
        x = x + _conv_block(sd, p, x, train, use_dropout, drop, True)
    return x


def gan_loss(pred, target_is_real):
    t = torch.ones_like(pred) if target_is_real else torch.zeros_like(pred)
    return F.binary_cross_entropy_with_logits(pred, t)


VGG_MEAN = (0.485, 0.456, 0.406)
VGG_STD = (0.229, 0.224, 0.225)


def vgg_features(vgg_sd, x, perceptual_layers=3):
    """VGG19.features[0 : perceptual_layers + 1] (L1_plus_perceptualLoss.py:22-27): conv3-64 (0), ReLU (1), conv64-64 (2),
    ReLU (3) -- the shipped value 3 gives all four; 0 / 2 end on a convolution, without its ReLU."""
    h = F.conv2d(x, vgg_sd["0.weight"], vgg_sd["0.bias"], padding=1)
    if perceptual_layers >= 1:
        h = F.relu(h)
    if perceptual_layers >= 2:
        h = F.conv2d(h, vgg_sd["2.weight"], vgg_sd["2.bias"], padding=1)
    if perceptual_layers >= 3:
        h = F.relu(h)
    assert perceptual_layers <= 3, "deeper slices (max-pool onwards) are not restated"
    return h


def l1_plus_perceptual(vgg_sd, inputs, targets, lambda_l1, lambda_perc, percep_is_l1=1, perceptual_layers=3):
    loss_l1 = F.l1_loss(inputs, targets) * lambda_l1
    mean = torch.tensor(VGG_MEAN, device=inputs.device).view(1, 3, 1, 1)
    std = torch.tensor(VGG_STD, device=inputs.device).view(1, 3, 1, 1)
    f = vgg_features(vgg_sd, ((inputs + 1) / 2 - mean) / std, perceptual_layers)
    t = vgg_features(vgg_sd, ((targets + 1) / 2 - mean) / std, perceptual_layers).detach()
    if percep_is_l1 == 1:
        loss_p = F.l1_loss(f, t) * lambda_perc
    else:
        loss_p = F.mse_loss(f, t) * lambda_perc
    return loss_l1 + loss_p, loss_l1, loss_p


class ImagePoolRef:
    """util/image_pool.py:7-34 (python ``random`` drives the swaps)."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.num_imgs = 0
        self.images = []

    def query(self, images):
        if self.pool_size == 0:
            return images
        out = []
        for image in images:
            image = image.unsqueeze(0)
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(image)
                out.append(image)
            else:
                if random.uniform(0, 1) > 0.5:
                    rid = random.randint(0, self.pool_size - 1)
                    tmp = self.images[rid].clone()
                    self.images[rid] = image
                    out.append(tmp)
                else:
                    out.append(image)
        return torch.cat(out, 0)


# ------------------------------------------------------------------------------------------------ trainer
class OracleTrainer:
    """One G + D_PP + D_PB step exactly in the order of MMHandModel.optimize_parameters (:310-330)."""

    def __init__(self, sd_g, sd_dpb, sd_dpp, sd_vgg, lambda_A=10.0, lambda_B=10.0, lambda_GAN=5.0, lr=2e-4,
                 beta1=0.5, pool_size=50, use_dropout_g=True, use_dropout_d=True, dropout="off", seed=49,
                 device="cpu", dg_ratio=1):
        def prep(sd):
            out = {}
            for k, v in sd.items():
                v = v.detach().clone().to(device)
                if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
                    v.requires_grad_(True)
                out[k] = v
            return out

        self.g, self.dpb, self.dpp = prep(sd_g), prep(sd_dpb), prep(sd_dpp)
        self.vgg = {k: v.detach().clone().to(device) for k, v in sd_vgg.items()}
        self.lA, self.lB, self.lG = lambda_A, lambda_B, lambda_GAN
        self.udg, self.udd = use_dropout_g, use_dropout_d
        self.dropout, self.seed, self.step_id = dropout, seed, 0
        self.dg_ratio = dg_ratio          # MMHandModel.py:320-329: each discriminator is stepped DG_ratio times
        params = lambda sd: [v for v in sd.values() if v.requires_grad]
        self.opt_g = torch.optim.Adam(params(self.g), lr=lr, betas=(beta1, 0.999))
        self.opt_dpb = torch.optim.Adam(params(self.dpb), lr=lr, betas=(beta1, 0.999))
        self.opt_dpp = torch.optim.Adam(params(self.dpp), lr=lr, betas=(beta1, 0.999))
        self.pool_pp, self.pool_pb = ImagePoolRef(pool_size), ImagePoolRef(pool_size)

    def _drop(self, net_id):
        return DropCtx(self.dropout, self.seed, self.step_id, net_id)

    def _d(self, sd, x, net_id):
        return discriminator_forward(sd, x, True, self.udd, drop=self._drop(net_id))

    def step(self, H1, P1, D1, H2, P2, D2):
        # net ids for dropout keys: 0 = G; D_PB: 1 (G step), 2 (real), 3 (fake); D_PP: 4, 5, 6
        fake = generator_forward(self.g, [H1, torch.cat((P1, P2), 1), torch.cat((D1, D2), 1)], True, self.udg,
                                 drop=self._drop(0))
        self.fake = fake
        # ---- G (backward_G)
        self.opt_g.zero_grad()
        l_pb = gan_loss(self._d(self.dpb, torch.cat((fake, P2), 1), 1), True)
        l_pp = gan_loss(self._d(self.dpp, torch.cat((fake, H1), 1), 4), True)
        l1tot, l1, lp = l1_plus_perceptual(self.vgg, fake, H2, self.lA, self.lB)
        pair_gan = (l_pb * self.lG + l_pp * self.lG) / 2
        (l1tot + pair_gan).backward()
        self.opt_g.step()
        # ---- D_PP (DG_ratio times, a fresh pool query each time)
        for _ in range(self.dg_ratio):
            self.opt_dpp.zero_grad()
            real = torch.cat((H2, H1), 1)
            fk = self.pool_pp.query(torch.cat((fake, H1), 1).detach())
            loss_dpp = (gan_loss(self._d(self.dpp, real, 5), True) * self.lG +
                        gan_loss(self._d(self.dpp, fk, 6), False) * self.lG) * 0.5
            loss_dpp.backward()
            self.opt_dpp.step()
        # ---- D_PB
        for _ in range(self.dg_ratio):
            self.opt_dpb.zero_grad()
            real = torch.cat((H2, P2), 1)
            fk = self.pool_pb.query(torch.cat((fake, P2), 1).detach())
            loss_dpb = (gan_loss(self._d(self.dpb, real, 2), True) * self.lG +
                        gan_loss(self._d(self.dpb, fk, 3), False) * self.lG) * 0.5
            loss_dpb.backward()
            self.opt_dpb.step()
        self.step_id += 1
        return {"pair_L1loss": l1tot.item(), "D_PP": loss_dpp.item(), "D_PB": loss_dpb.item(),
                "pair_GANloss": pair_gan.item(), "origin_L1": l1.item(), "perceptual": lp.item()}
