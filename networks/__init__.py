"""Drop-in package for the benchmark harness of the reference (baselines/quantitative_on_benchmarks): this repository's
``networks.model_variants`` shadows the reference's; its other ``networks.*`` modules stay importable when that
directory follows this repository on sys.path (same arrangement as ``util`` and ``data``)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
