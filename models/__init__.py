"""Drop-in package: the modules of this repository shadow the reference's same-named ones; every other module of the
reference's package of the same name (e.g. util.visualizer, util.util, models.utils) stays importable when the reference
tree follows this repository on sys.path -- the package path is extended over sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
