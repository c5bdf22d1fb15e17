"""Drop-in package: ``data.mmhand_dataset_data_loader`` of this repository shadows the reference's (device-side input
pipeline); every other ``data.*`` module of the reference stays importable when the reference tree follows this
repository on sys.path -- the package path is extended over sys.path (same arrangement as ``util``)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
