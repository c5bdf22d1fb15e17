"""mmhand_b200: B200-native (sm_100a) kernels and engines behind the MM-HAND generator / discriminator / loss
hot path. The reference-facing modules live at the repository root (models/, losses/, util/), exactly where the
reference's train.py and aug.py import them from."""
__version__ = "0.1.0"
