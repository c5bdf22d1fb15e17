"""Checkpoint ABI: the drop-in modules expose exactly the reference's state_dict keys, shapes and dtypes
(golden state_dicts were written by the reference's own modules, oracle/make_golden.py)."""
import os

import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_generator_and_discriminator_keys_match_reference():
    from models.Discriminator import Discriminator
    from models.Generator import Generator
    from models.network_utils import get_norm_layer
    gold = torch.load(os.path.join(GOLD, "nets_ngf4.pt"))
    norm = get_norm_layer('batch')
    for mine, ref in ((Generator([3, 42, 6], 3, 4, norm, True, 9), gold["g_sd"]),
                      (Generator([3, 42, 6], 3, 4, norm, False, 9), gold["g2_sd"]),
                      (Discriminator(24, 4, norm, True, 3, [], 'reflect', False, 2), gold["d_sd"]),
                      (Discriminator(24, 4, norm, False, 3, [], 'reflect', False, 2), gold["d2_sd"])):
        sd = mine.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        for k in sd:
            assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype, k
        mine.load_state_dict(ref)          # a reference checkpoint loads as is


def test_full_size_parameter_counts():
    from models.Discriminator import Discriminator
    from models.Generator import Generator
    from models.network_utils import get_norm_layer
    norm = get_norm_layer('batch')
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(Generator([3, 42, 6], 3, 64, norm, True, 9)) == 71272835
    assert n(Discriminator(24, 64, norm, True, 3)) == 3986816
    assert n(Discriminator(6, 64, norm, True, 3)) == 3930368


def test_instance_norm_models_are_accepted_for_checkpoint_compatibility():
    """norm='instance' (network_utils.get_norm_layer: InstanceNorm2d(affine=False) => biased convs, no norm entries): the
    constructors accept it and expose the reference's state_dict (tests/golden/nets_instance_ngf4.pt, written by the
    reference's own modules); the B200 path computes batch norm only, so forward refuses."""
    import pytest
    from models.Discriminator import Discriminator
    from models.Generator import Generator
    from models.network_utils import get_norm_layer
    gold = torch.load(os.path.join(GOLD, "nets_instance_ngf4.pt"))
    norm = get_norm_layer('instance')
    g = Generator([3, 42, 6], 3, 4, norm, True, 9)
    d = Discriminator(24, 4, norm, True, 3, [], 'reflect', False, 2)
    for mine, ref in ((g, gold["g_in_sd"]), (d, gold["d_in_sd"])):
        sd = mine.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        for k in sd:
            assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype, k
        mine.load_state_dict(ref)
    with pytest.raises(NotImplementedError):
        g([torch.zeros(1, 3, 32, 32), torch.zeros(1, 42, 32, 32), torch.zeros(1, 6, 32, 32)])
    with pytest.raises(NotImplementedError):
        d(torch.zeros(1, 24, 32, 32))
