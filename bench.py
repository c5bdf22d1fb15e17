"""bench.py -- G+D train-step images/sec @256x256 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU arithmetic (oracle port) on host cores

A step = MMHandModel.optimize_parameters() on one synthetic batch (configs[2]: full G+D training with
L1 + VGG19 perceptual loss, batch 16/GPU, 256x256, BN, dropout on). ``value`` times K steps with the batch resident
in HBM (CUDA events, barrier + synchronize on both sides, max over ranks); ``e2e`` times K steps through the public
API from pinned host memory (H2D of the six input tensors + D2H of the six loss scalars inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# there is no network and no pretrained VGG19 file on the GPU boxes: the perceptual loss runs on seeded random-init
# VGG19[:4] features (same architecture and FLOPs; recorded in the JSON line's config)
os.environ.setdefault("MMH_VGG19_RANDOM", "1")

GFLOP_PER_IMG = 2490.0       # minimal required G+D train step, 2*MACs, un-padded channels


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 20 (train, infer), 245 (raster: ~1M poses)")
    ap.add_argument("--workload", default="train", choices=["train", "infer", "raster", "jointsmap"],
                    help="train = configs[2] (the headline metric); infer = configs[1] (generator-only inference, "
                         "batch 32); raster = configs[3] (keypoint -> heatmap rasteriser, 4096 poses per step)")
    ap.add_argument("--poses-per-step", type=int, default=4096)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: 16 train, 32 infer)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="train workload on one GPU: skip the short infer / raster runs reported under 'secondary'")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 245 if a.workload == "raster" else (25 if a.workload == "jointsmap" else 20)
    if a.batch is None:
        a.batch = 32 if (a.workload == "infer" and a.impl == "ours") else 16
    return a


def synth_batch(B, S, seed, pin=False):
    """RHD/STB-shaped synthetic sample dict in the reference loader's fp32 form: images U(-1,1), sparse
    heatmaps in [0,1], depth x3."""
    import torch
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    d1, d2 = r(B, 1, S, S) * 2 - 1, r(B, 1, S, S) * 2 - 1
    out = dict(H1=r(B, 3, S, S) * 2 - 1, P1=(r(B, 21, S, S) > 0.984).float() * r(B, 21, S, S),
               D1=d1.expand(B, 3, S, S).contiguous(), H2=r(B, 3, S, S) * 2 - 1,
               P2=(r(B, 21, S, S) > 0.984).float() * r(B, 21, S, S), D2=d2.expand(B, 3, S, S).contiguous())
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def synth_compact_batch(B, S, seed, pin=True):
    """The same kind of sample as the device-side input pipeline delivers it (mmhand_b200/loader.py): uint8 colour and
    depth frames as cv2.imread returns them (BGR; depth = 256*G + R in [200, 700) mm) and float64 keypoints inside the
    frame. MMHandModel.set_input turns them into the six fp32 tensors on the device (21 Gaussian heatmaps per pose,
    ~98.4 % zeros: the real distribution)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in ("H1", "H2"):
        out[k + "_u8"] = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8)
    for k in ("D1", "D2"):
        d = torch.randint(200, 700, (B, S, S), generator=g, dtype=torch.int32)
        f = torch.zeros(B, S, S, 3, dtype=torch.uint8)
        f[..., 1] = (d // 256).to(torch.uint8)
        f[..., 2] = (d % 256).to(torch.uint8)
        out[k + "_u8"] = f
    for k in ("P1", "P2"):
        out[k + "_uv"] = (torch.rand(B, 21, 2, generator=g, dtype=torch.float64) * (S - 32) + 16)
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    out["u8_bgr"] = True
    return out


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in o.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.startswith("Active")})
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=0)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_reference_rate(S, seconds, max_steps=None, B=1):
    """Reference arithmetic (oracle port of MMHandModel.optimize_parameters) on the host cores, fp32, batch B.
    Returns (images per second, timed steps, threads)."""
    import random

    import torch
    import torchvision

    from oracle import patn_ref as O
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(49)
    random.seed(49)

    def init(sd):
        for k, v in sd.items():
            if k.endswith("weight") and v.dim() == 4:
                v.normal_(0.0, 0.02)
            elif k.endswith("weight") and v.dim() == 1:
                v.normal_(1.0, 0.02)
        return sd

    # state_dicts with the reference's key names and shapes, built without touching the CUDA engines
    from models.Discriminator import Discriminator
    from models.Generator import Generator
    from models.network_utils import get_norm_layer
    norm = get_norm_layer('batch')
    g = init({k: v.clone() for k, v in Generator([3, 42, 6], 3, 64, norm, True, 9).state_dict().items()})
    dpb = init({k: v.clone() for k, v in Discriminator(24, 64, norm, True, 3).state_dict().items()})
    dpp = init({k: v.clone() for k, v in Discriminator(6, 64, norm, True, 3).state_dict().items()})
    vgg = torchvision.models.vgg19(weights=None).features[:4].state_dict()
    tr = O.OracleTrainer(g, dpb, dpp, vgg, dropout="hash", seed=49, device="cpu")
    b = synth_batch(B, S, 7)
    t_w = time.time()
    tr.step(b["H1"], b["P1"], b["D1"], b["H2"], b["P2"], b["D2"])       # warm-up
    t_w = time.time() - t_w
    n, t0 = 0, time.time()
    while True:
        tr.step(b["H1"], b["P1"], b["D1"], b["H2"], b["P2"], b["D2"])
        n += 1
        el = time.time() - t0
        if el + t_w >= seconds or (max_steps and n >= max_steps):      # never start a step that overruns the budget
            break
    return B * n / el, n, torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = 150.0
    rate, n, cores = cpu_reference_rate(a.size, budget, max_steps=max(1, a.steps), B=a.batch)
    line = {
        "impl": "reference", "metric": "G+D train-step images/sec @256x256", "value": rate, "unit": "images/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 / rate,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: full G+D training, L1+VGG19 perceptual loss, 256x256, BN, dropout on",
                   "per_gpu_batch": a.batch, "global_batch": a.batch, "frame": a.size,
                   "parallelism": "host cores (rank 0 only)"},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d timed step(s) of batch %d after one warm-up step, inside a 150 s budget (reference "
                                   "arithmetic restated in oracle/patn_ref.py, torch fp32 CPU ops, all host threads; the "
                                   "reference is Python + apex and cannot travel to the box)" % (n, a.batch)},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(a):
    import random

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", init_method="env://", device_id=torch.device("cuda", local))
    from mmhand_b200 import runtime
    from models.MMHandModel import MMHandModel
    from mmhand_b200.options import make_opt

    torch.manual_seed(49)
    random.seed(49 + rank)
    B, S = a.batch, a.size
    opt = make_opt(batchSize=B, fineSize=S, local_rank=local, gpu=local, seed=49, distributed=(world > 1))
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = MMHandModel(opt)
    ops = runtime.get_ops(torch.device("cuda", local))
    # e2e feed: what the device-side input pipeline delivers (uint8 frames + keypoints, pinned); device-resident feed:
    # the six fp32 tensors set_input makes of those very batches (already in HBM when the timed region starts)
    host = [synth_compact_batch(B, S, 1000 + 17 * rank + i, pin=True) for i in range(2)]
    dev = []
    for h in host:
        model.set_input(h)
        dev.append({k: getattr(model, "input_" + k).clone() for k in ("H1", "P1", "D1", "H2", "P2", "D2")})
    torch.cuda.synchronize()
    model._in_shapes = None          # the timed feeds re-create the static input buffers once, in warm-up
    h2d = sum(v.numel() * v.element_size() for v in host[0].values() if isinstance(v, torch.Tensor))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, feed, read_back):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launches
        e0.record()
        for i in range(n):
            model.set_input(feed[i % 2])
            model.optimize_parameters()
            if read_back:
                errs = torch.stack([v.reshape(()) for v in model.get_current_errors().values()]).cpu()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, ops.launches - l0

    # warm-up: at least 3 steps, and enough to fill the image pools (pool_size / B steps) so that the timed steps
    # run the steady-state pool path (swaps) like any step of a real epoch
    warm = max(a.warmup, 3, -(-opt.pool_size // B) + 1)
    for i in range(warm):
        model.set_input(dev[i % 2])
        model.optimize_parameters()
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(a.steps, dev, False)
    for i in range(2):                                   # the compact feed's staging buffers / first launches
        model.set_input(host[i])
        model.optimize_parameters()
    ms_e2e, _ = timed(a.steps, host, True)
    sampler.stop_flag = True
    sampler.join(timeout=3)
    value = B * world * a.steps / (ms / 1000.0)
    e2e = B * world * a.steps / (ms_e2e / 1000.0)

    # rooflines: one extra instrumented (eager, single-stream) step with a CUDA-event pair around every launch of the
    # tensor-core kernels on the launching stream
    recs = {"conv": [], "wgrad": []}

    def hook(kind, tag, plan, launch):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        d = plan.desc
        if kind == "conv":
            fl = 2.0 * d.M * d.N * d.C * d.T * (d.Hv * d.Wv) / float(d.Hg * d.Wg)
        else:
            fl = 2.0 * d.M * d.N_store * d.C_store * d.T
        recs[kind].append((e0, e1, fl))

    ew = {}

    def ew_hook(kind, nbytes, launch):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        ew.setdefault(kind, []).append((e0, e1, float(nbytes)))

    model.use_tape = False
    ops.conv_hook, ops.ew_hook = hook, ew_hook
    side, ops.side_stream = ops.side_stream, None      # serial launches: a kernel's events bracket that kernel alone
    chains, ops.chains = ops.chains, [None]
    model.set_input(dev[0])
    model.optimize_parameters()
    torch.cuda.synchronize()
    ops.side_stream, ops.chains = side, chains
    ops.conv_hook = ops.ew_hook = None
    model.use_tape = True
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    src = "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback 1.4 PF sustained"

    def roof(kind, kernel, traffic, note):
        t = sum(e0.elapsed_time(e1) for e0, e1, _ in recs[kind]) / 1000.0
        f = sum(f for _, _, f in recs[kind])
        ach = f / t / 1e12 if t > 0 else 0.0
        return {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": src, "launches_per_step": len(recs[kind]),
                "kernel_ms_per_step": t * 1000.0, "flops_note": note}

    roofline = roof("conv", "conv2_kernel (tcgen05 implicit-GEMM fprop/dgrad, csrc/tc_conv2.cu)",
                    # dram__bytes_read.sum + dram__bytes_write.sum of one 3x3 512->512 launch at batch 16 (ncu --set full,
                    # profiles/r01_conv2_ncu_full_v6.txt): 76.3 MB read + 36.3 MB written, against 71.4 (activations) +
                    # 4.7 (weights) + 71.4 (output) MB algorithmic -- the output stays in L2 for its consumer
                    {"bytes_per_launch_3x3_512": 112.6e6, "algorithmic_bytes": 147.5e6} if (B == 16 and S == 256) else None,
                    "algorithmic 2*MACs of the valid output positions (padded-grid rows excluded); the stems' channel "
                    "padding (3/6/24/42 -> 16/16/32/48) is included")
    roofline["step_tensor_util"] = GFLOP_PER_IMG * 1e9 * value / world / (peak * 1e12)
    roofline_wgrad = roof("wgrad", "wgrad2_kernel (tcgen05 weight gradient, csrc/tc_wgrad2.cu)",
                          # one 3x3 256->256 launch at batch 16 (ncu --set full, profiles/r02_wgrad2_ncu_full.txt): 74.3 MB read +
                          # 4.3 MB written against 35.7 (dY) + 35.7 (activations) + 2.4 (dw) MB algorithmic
                          {"bytes_per_launch_3x3_256": 78.5e6, "algorithmic_bytes": 73.8e6} if (B == 16 and S == 256) else None,
                          "algorithmic 2*MACs over all grid rows with un-padded channel counts")
    hbm = peaks.get("hbm_gbs", 6400.0)
    names = {"bn_bwd": "rows_reduce_fin_kernel<BnLeanReduceF> + rows_pg_kernel<BnLeanApplyF> (BatchNorm backward: sums, "
                       "then data gradient; csrc/elementwise.cu)",
             "norm_act": "pg_kernel<NormActF> (BN apply + ReLU + dropout + residual + next layer's halo)"}
    roofline_ew = []
    for kind in ("bn_bwd", "norm_act"):
        r = ew.get(kind, [])
        t = sum(e0.elapsed_time(e1) for e0, e1, _ in r) / 1000.0
        by = sum(b for _, _, b in r)
        ach = by / t / 1e9 if t > 0 else 0.0
        roofline_ew.append({"bound": "hbm", "kernel": names[kind], "achieved": ach, "peak": hbm, "unit": "GB/s",
                            # one C = 256, 64 x 64 pair at batch 16 (profiles/r02_bn_lean_ncu_full.txt): reduce 69.4 MB read;
                            # apply 69.3 MB read + 33.6 MB written (8.2 MB of it reach DRAM inside the launch)
                            "traffic": ({"bytes_per_pair_c256": 69.4e6 + 69.3e6 + 33.6e6, "algorithmic_bytes": 67.1e6 * 2 + 33.6e6}
                                        if (kind == "bn_bwd" and B == 16 and S == 256) else None),
                            "frac": ach / hbm, "launches_per_step": len(r),
                            "kernel_ms_per_step": t * 1000.0,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback",
                            "bytes_note": "algorithmic bytes: bf16 elements read and written once (reduce 4 B, apply 6 B, "
                                          "norm_act 2 + 2 B incl. the halo, + fp32 residual / trunk where present); "
                                          "35-140 MB per launch"})
    chains_used = len(model.netG.engine(B, S, S, model.world)._chains())
    if world > 1:
        dist.barrier()                       # every rank is done measuring before the group goes away
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "G+D train-step images/sec @256x256", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": a.steps, "warmup": warm, "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[2]: full G+D training, L1+VGG19 perceptual loss, 256x256, BN, dropout on",
                   "per_gpu_batch": B, "global_batch": B * world, "frame": S, "parallelism": "dp%d" % world,
                   "l2": "per-step working set (activations > 5 GB) far exceeds the 126 MB L2",
                   "vgg_weights": "random-init (no network)",
                   "syncbn": ("peer" if (model.world is not None and model.world.peer is not None) else
                              ("nccl" if world > 1 else "local")),
                   "pdl": bool(ops.lib.mmh_get_pdl()),
                   "g_update_stream": os.environ.get("MMH_G_UPDATE_STREAM", "0" if world > 1 else "1") != "0",
                   "grad_allreduce": getattr(model, "grad_sync_mode", "none") if world > 1 else "none",
                   "bn_bwd_in_dgrad_epilogue": os.environ.get("MMH_FUSE_BN_BWD", "0") != "0",
                   "layer_chain_streams": chains_used,
                   "e2e_feed": "uint8 frames + float64 keypoints from pinned host memory (mmhand_b200/loader.py form); "
                               "heatmaps rasterised and frames normalised on the device inside set_input"},
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 6 * 4,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline,
        "rooflines_other": [roofline_wgrad] + roofline_ew,
    }
    if world == 1 and not a.no_cpu_baseline:
        rate, n, cores = cpu_reference_rate(S, a.cpu_seconds)
        line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": "%d step(s) of batch 1, oracle port of the reference step, torch fp32 CPU" % n}
    return line


# ======================================================================================= secondary workloads
def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", init_method="env://", device_id=torch.device("cuda", local))
    return world, rank, local


def _timed_loop(n, body, world):
    """K calls of body(i) between barrier + synchronize, CUDA events on the launching stream, max over ranks."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        body(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def run_infer(a):
    """configs[1]: Generator-only inference (the aug.py path), batch 32, 256x256, eval-mode BN, no_grad.
    Shards by image across ranks, no collective."""
    import torch

    world, rank, local = _dist_setup()
    from mmhand_b200 import runtime
    from models.Generator import Generator
    from models.network_utils import get_norm_layer, init_weights
    B, S = a.batch, a.size
    torch.manual_seed(49)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        g = Generator([3, 42, 6], 3, 64, get_norm_layer('batch'), True, 9).to(torch.device("cuda", local))
        init_weights(g, 'normal')
    g.eval()
    ops = runtime.get_ops(torch.device("cuda", local))
    hosts, devs = [], []
    for i in range(2):
        b = synth_batch(B, S, 2000 + 17 * rank + i, pin=True)
        h = [b["H1"], torch.cat((b["P1"], b["P2"]), 1).pin_memory(), torch.cat((b["D1"], b["D2"]), 1).pin_memory()]
        hosts.append(h)
        devs.append([t.cuda(non_blocking=True) for t in h])
    # end to end = the public aug.py path of this repository (mmhand_b200.augment.generate_batch): the loader's compact
    # batch (uint8 frames + float64 keypoints, pinned) -> six tensors on the device -> generator -> BGR uint8 images
    # (the bytes cv2.imwrite stores) -> pinned host buffer
    from mmhand_b200 import augment
    from mmhand_b200.loader import CompactBatch
    compact = [CompactBatch(synth_compact_batch(B, S, 2100 + 17 * rank + i, pin=True)) for i in range(2)]
    out_host = torch.empty(B, S, S, 3, dtype=torch.uint8).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in compact[0].values() if torch.is_tensor(v))

    def dev_step(i):
        with torch.no_grad():
            g(devs[i % 2])

    def e2e_step(i):
        # a fresh CompactBatch per step: the device tensors it materialises are not cached across steps
        augment.generate_batch(g, CompactBatch(compact[i % 2]), host_out=out_host)
        torch.cuda.current_stream().synchronize()       # aug.py writes each batch of images to disk

    warm = max(a.warmup, 3)
    for i in range(warm):
        dev_step(i)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.launches
    ms = _timed_loop(a.steps, dev_step, world)
    launches = ops.launches - l0
    ms_e2e = _timed_loop(a.steps, e2e_step, world)
    sampler.stop_flag = True
    sampler.join(timeout=3)
    value = B * world * a.steps / (ms / 1000.0)
    e2e = B * world * a.steps / (ms_e2e / 1000.0)

    recs = []

    def hook(kind, tag, plan, launch):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        d = plan.desc
        recs.append((e0, e1, 2.0 * d.M * d.N * d.C * d.T * (d.Hv * d.Wv) / float(d.Hg * d.Wg)))

    ops.conv_hook = hook
    chains, ops.chains = ops.chains, [None]       # serial launches on one stream: a kernel's events bracket that kernel
    os.environ["MMH_INFER_TAPE"] = "0"            # alone (torch events see only torch's current stream); eager launches
    g._engines.clear()                            # (the hook wraps eager conv launches), on a fresh engine
    dev_step(0)
    torch.cuda.synchronize()
    os.environ.pop("MMH_INFER_TAPE", None)
    ops.chains = chains
    ops.conv_hook = None
    t_conv = sum(e0.elapsed_time(e1) for e0, e1, _ in recs) / 1000.0
    f_conv = sum(f for _, _, f in recs)
    peaks = _peaks()
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = f_conv / t_conv / 1e12 if t_conv > 0 else 0.0
    if rank != 0:
        return
    gflop_img = 611.68
    line = {
        "metric": "generator inference images/sec @256x256", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": a.steps, "warmup": warm, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[1]: Generator-only inference (aug.py path), eval-mode BN, no_grad",
                   "per_gpu_batch": B, "frame": S, "parallelism": "dp%d (independent images, no collective)" % world,
                   "l2": "activations of one batch (> 2 GB) exceed the 126 MB L2"},
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": out_host.numel(), "ms_per_step": ms_e2e / a.steps,
                "path": "mmhand_b200.augment.generate_batch on the loader's compact batch; BGR uint8 images back"},
        "gpu_launches": launches, "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "kernel": "conv2_kernel (tcgen05 implicit-GEMM fprop, csrc/tc_conv2.cu)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback 1.4 PF sustained",
                     "launches_per_step": len(recs), "kernel_ms_per_step": t_conv * 1000.0,
                     "step_tensor_util": gflop_img * 1e9 * value / world / (peak * 1e12)},
    }
    if world == 1 and not a.no_cpu_baseline:
        from oracle import patn_ref as O
        torch.set_num_threads(os.cpu_count() or 1)
        sd = {k: v.detach().cpu().clone() for k, v in g.state_dict().items()}
        x = [t[:1].clone() for t in hosts[0]]
        with torch.no_grad():
            O.generator_forward(sd, x, train=False)
            n, t0 = 0, time.time()
            while time.time() - t0 < a.cpu_seconds:
                O.generator_forward(sd, x, train=False)
                n += 1
        rate = n / (time.time() - t0)
        line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "%d eval forwards of batch 1, oracle port of Generator.forward, torch fp32 CPU" % n}
    return line


def run_raster(a):
    """configs[3]: keypoint -> 21-joint Gaussian heatmaps (mmh_heatmap_rasterize), K steps of P poses each
    (default 245 x 4096 = 1,003,520 poses). Poses shard across ranks, no collective. Output buffers alternate
    between two P-pose blocks (2 x 22.5 GB at P = 4096), far larger than L2."""
    import numpy as np
    import torch

    world, rank, local = _dist_setup()
    from mmhand_b200 import runtime
    from mmhand_b200.rasterize import get_heatmaps
    P, S = a.poses_per_step, a.size
    dev = torch.device("cuda", local)
    ops = runtime.get_ops(dev)
    rng = np.random.RandomState(49 + rank)
    uv_host = torch.from_numpy(rng.uniform(16.0, 240.0 * S / 256.0, size=(2, P, 21, 2))).pin_memory()
    uv_dev = uv_host.to(dev)
    outs = [torch.empty(P, 21, S, S, dtype=torch.float32, device=dev) for _ in range(2)]
    chk = torch.zeros(1, dtype=torch.float64, device=dev)
    chk_host = torch.zeros(1, dtype=torch.float64).pin_memory()

    def dev_step(i):
        get_heatmaps(uv_dev[i % 2], (S, S), out=outs[i % 2])

    def e2e_step(i):
        uv = uv_host[i % 2].to(dev, non_blocking=True)
        o = get_heatmaps(uv, (S, S), out=outs[i % 2])
        chk_host.copy_(o[:, :, S // 2, :].sum(dtype=torch.float64).reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    warm = max(a.warmup, 3)
    for i in range(warm):
        dev_step(i)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.launches
    ms = _timed_loop(a.steps, dev_step, world)
    launches = ops.launches - l0
    ms_e2e = _timed_loop(a.steps, e2e_step, world)
    sampler.stop_flag = True
    sampler.join(timeout=3)
    value = P * world * a.steps / (ms / 1000.0)
    e2e = P * world * a.steps / (ms_e2e / 1000.0)
    # one kernel per step: its average launch duration is the step time
    bytes_per_pose = 21 * S * S * 4 + 21 * 2 * 8
    peaks = _peaks()
    peak = peaks.get("hbm_gbs", 6500.0)
    achieved = bytes_per_pose * P * a.steps / (ms / 1000.0) / 1e9
    if rank != 0:
        return
    line = {
        "metric": "keypoint->heatmap rasterisation poses/sec @256x256x21", "value": value, "unit": "poses/s",
        "n_gpus": world, "steps": a.steps, "warmup": warm, "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[3]: 21-joint Gaussian heatmaps, fp64 arithmetic, fp32 maps",
                   "poses_per_step": P, "poses_total": P * a.steps * world, "frame": S,
                   "parallelism": "dp%d (independent poses, no collective)" % world,
                   "l2": "two alternating %.1f GB output blocks, far larger than L2" % (bytes_per_pose * P / 1e9)},
        "e2e": {"value": e2e, "unit": "poses/s", "h2d_bytes_per_step": P * 21 * 2 * 8, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / a.steps,
                "note": "maps stay in HBM for the training step that consumes them; the host reads one checksum"},
        "gpu_launches": launches, "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": "raster_kernel (csrc/raster.cu)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one 4096-pose launch, ncu --set full
                     # (profiles/r01_raster_ncu_full_v6.txt): 5.37 MB read + 22.4897 GB written
                     "traffic": 22495045672 if (P == 4096 and S == 256) else None,
                     "algorithmic_bytes_per_launch": bytes_per_pose * P,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback",
                     "bytes_per_pose": bytes_per_pose},
    }
    if world == 1 and not a.no_cpu_baseline:
        from oracle.raster_ref import get_heatmaps as ref_heatmaps
        uvn = uv_host[0].numpy()
        n, t0 = 0, time.time()
        while time.time() - t0 < min(a.cpu_seconds, 15.0):
            ref_heatmaps(uvn[n % P], (S, S))
            n += 1
        rate = n / (time.time() - t0)
        line["cpu_baseline"] = {"value": rate, "unit": "poses/s", "cores": 1, "kind": "port",
                                "sample": "%d poses, numpy restatement of Genericdataset.get_heatmaps (one DataLoader worker)" % n}
    return line


def run_jointsmap(a):
    """configs[3], second half: depth-ordered part maps (mmh_jointsmap_rasterize), K steps of P poses, compact uint8
    maps. Integer scan conversion per bone: bounded by instruction issue / latency, not by HBM (64 KB per pose)."""
    import numpy as np
    import torch

    world, rank, local = _dist_setup()
    from mmhand_b200 import runtime
    from mmhand_b200.rasterize import generate_jointsmap
    P, S = a.poses_per_step, a.size
    dev = torch.device("cuda", local)
    ops = runtime.get_ops(dev)
    rng = np.random.RandomState(49 + rank)
    uv_host = torch.from_numpy(rng.uniform(16.0, 240.0 * S / 256.0, size=(2, P, 21, 2))).pin_memory()
    z_host = torch.from_numpy(rng.uniform(200.0, 700.0, size=(2, P, 21))).pin_memory()
    uv_dev, z_dev = uv_host.to(dev), z_host.to(dev)
    chk_host = torch.zeros(1, dtype=torch.float64).pin_memory()

    def dev_step(i):
        return generate_jointsmap(uv_dev[i % 2], z_dev[i % 2], S, S, dtype=torch.uint8)

    def e2e_step(i):
        o = generate_jointsmap(uv_host[i % 2].to(dev, non_blocking=True), z_host[i % 2].to(dev, non_blocking=True), S, S,
                               dtype=torch.uint8)
        chk_host.copy_(o[:, S // 2, :].sum(dtype=torch.float64).reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    warm = max(a.warmup, 3)
    for i in range(warm):
        dev_step(i)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.launches
    ms = _timed_loop(a.steps, dev_step, world)
    launches = ops.launches - l0
    ms_e2e = _timed_loop(a.steps, e2e_step, world)
    sampler.stop_flag = True
    sampler.join(timeout=3)
    value = P * world * a.steps / (ms / 1000.0)
    e2e = P * world * a.steps / (ms_e2e / 1000.0)
    bytes_per_pose = S * S + 21 * 3 * 8
    peaks = _peaks()
    peak = peaks.get("hbm_gbs", 6500.0)
    achieved = bytes_per_pose * P * a.steps / (ms / 1000.0) / 1e9
    if rank != 0:
        return
    line = {
        "metric": "joints->part-map rasterisation poses/sec @256x256", "value": value, "unit": "poses/s",
        "n_gpus": world, "steps": a.steps, "warmup": warm, "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": {"workload": "configs[3]: depth-ordered part maps (generate_jointsmap), 20 bones, uint8 maps",
                   "poses_per_step": P, "poses_total": P * a.steps * world, "frame": S,
                   "parallelism": "dp%d (independent poses, no collective)" % world,
                   "l2": "outputs of consecutive steps alternate; 64 KB per pose"},
        "e2e": {"value": e2e, "unit": "poses/s", "h2d_bytes_per_step": P * 21 * 3 * 8, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": launches, "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": "jointsmap_kernel (csrc/jointsmap.cu)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "note": "integer scan conversion: issue/latency-bound, far below the HBM roofline by nature",
                     "bytes_per_pose": bytes_per_pose},
    }
    if world == 1 and not a.no_cpu_baseline:
        from oracle.jointsmap_ref import generate_jointsmap_cv2
        try:
            import cv2  # noqa: F401
            ref, kind = generate_jointsmap_cv2, "reference lines on the real cv2"
        except Exception:
            from oracle.jointsmap_ref import generate_jointsmap as ref
            kind = "pure-Python restatement (cv2 not importable)"
        uvn, zn = uv_host[0].numpy(), z_host[0].numpy()
        n, t0 = 0, time.time()
        while time.time() - t0 < min(a.cpu_seconds, 10.0):
            ref(uvn[n % P], zn[n % P], S, S)
            n += 1
        rate = n / (time.time() - t0)
        line["cpu_baseline"] = {"value": rate, "unit": "poses/s", "cores": 1, "kind": "port",
                                "sample": "%d poses, generate_jointsmap (%s), one DataLoader worker" % (n, kind)}
    return line


def _watchdog(seconds):
    """Multi-GPU runs only: a rank that is still running after `seconds` prints what it knows and exits, so that a
    collective that never completes costs a bounded time instead of the caller's whole time limit."""
    def fire():
        rank = int(os.environ.get("RANK", "0"))
        sys.stderr.write("bench.py watchdog: rank %d still running after %d s -- giving up\n" % (rank, seconds))
        try:                                   # where every thread of this rank is stuck
            import faulthandler
            faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        except Exception:
            pass
        sys.stderr.flush()
        if rank == 0:
            print(json.dumps({"error": "watchdog: multi-GPU run exceeded %d s" % seconds,
                              "n_gpus": int(os.environ.get("WORLD_SIZE", "1"))}), flush=True)
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


if __name__ == "__main__":
    args = parse()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        _watchdog(int(os.environ.get("MMH_BENCH_WATCHDOG_S", "300")))
    if args.impl == "reference":
        run_reference(args)
        sys.exit(0)
    fn = {"infer": run_infer, "raster": run_raster, "jointsmap": run_jointsmap, "train": run_ours}[args.workload]
    line = fn(args)
    if line is not None and args.workload == "train" and line["n_gpus"] == 1 and not args.no_secondary:
        # the other BASELINE configs, short runs, so that the driver's record carries them next to the headline
        import copy
        sec = {}
        for name, f, steps, batch in (("infer_configs1", run_infer, 10, 32), ("raster_configs3", run_raster, 60, None)):
            b = copy.copy(args)
            b.steps, b.warmup, b.no_cpu_baseline, b.workload = steps, 3, True, name.split("_")[0]
            if batch:
                b.batch = batch
            try:
                r = f(b)
                sec[name] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "roofline", "config")}
            except Exception as e:          # a failed side measurement must not cost the headline
                sec[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        line["secondary"] = sec
    if line is not None:
        print(json.dumps(line))
