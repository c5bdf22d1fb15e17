"""Generate tests/golden/patn2_ngf4.pt from the REAL reference modules of the benchmark harness (build container only):
the two-stream ``PATNetwork`` (baselines/quantitative_on_benchmarks/networks/model_variants.py) in eval and train mode,
and ``pytorch_ssim.ssim`` (baselines/quantitative_on_benchmarks/pytorch_ssim/__init__.py).

    python oracle/make_golden_variants.py
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("MMH_REFERENCE_ROOT", "/root/reference")
BENCH = os.path.join(REF, "baselines", "quantitative_on_benchmarks")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    mv = _load(os.path.join(BENCH, "networks", "model_variants.py"), "ref_model_variants")
    ss = _load(os.path.join(BENCH, "pytorch_ssim", "__init__.py"), "ref_pytorch_ssim")
    from oracle import ref_shims
    _, _, nu, _, _ = ref_shims.load_reference_nets()
    gen = torch.Generator().manual_seed(51)
    torch.manual_seed(51)
    S, B, ngf = 32, 2, 4
    norm = nu.get_norm_layer('batch')
    x = [torch.rand(B, 3, S, S, generator=gen) * 2 - 1, torch.rand(B, 3, S, S, generator=gen)]
    out = {"ngf": ngf, "x": x}
    g = mv.PATNetwork([3, 3], 3, ngf, norm, True, 9)
    nu.init_weights(g, 'normal')
    out["sd"] = {k: v.clone() for k, v in g.state_dict().items()}
    g.eval()
    with torch.no_grad():
        out["eval"] = g(x)
    g2 = mv.PATNetwork([3, 3], 3, ngf, norm, False, 9)          # no dropout: train mode without torch's RNG
    nu.init_weights(g2, 'normal')
    out["sd2"] = {k: v.clone() for k, v in g2.state_dict().items()}
    g2.train()
    with torch.no_grad():
        out["train"] = g2(x)
    out["sd2_after"] = {k: v.clone() for k, v in g2.state_dict().items() if "running" in k}
    a = torch.rand(2, 3, 40, 48, generator=gen)
    b = (a + 0.1 * torch.randn(2, 3, 40, 48, generator=gen)).clamp(0, 1)
    out["ssim_a"], out["ssim_b"] = a, b
    out["ssim_mean"] = ss.ssim(a, b)
    out["ssim_per_image"] = ss.ssim(a, b, size_average=False)
    out["ssim_same"] = ss.ssim(a, a)
    dst = os.path.join(ROOT, "tests", "golden", "patn2_ngf4.pt")
    torch.save(out, dst)
    print(dst, os.path.getsize(dst), float(out["ssim_mean"]), out["ssim_per_image"])


if __name__ == "__main__":
    main()
