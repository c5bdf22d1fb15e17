"""Hand-built option namespace with the reference's defaults (options/base_options.py:15-163,
options/train_options.py:7-35, scripts/mm-train-ratio.sh:22-39): what ``TrainOptions().parse()`` would return.
The reference's own parser crashes without ``--distributed``, so benches, smoke and tests build
``opt`` with this helper."""
import types


def make_opt(**over):
    """Keyword overrides on top of the shipped training defaults."""
    o = dict(batchSize=1, fineSize=256, H_input_nc=3, P_input_nc=21, D_input_nc=3, output_nc=3, ngf=64, ndf=64,
             n_layers_D=3, norm='batch', no_dropout=False, no_dropout_D=False, init_type='normal',
             G_n_downsampling=2, D_n_downsampling=2, padding_type='reflect', no_lsgan=True, lambda_A=10.0,
             lambda_B=10.0, lambda_GAN=5.0, L1_type='l1_plus_perL1', perceptual_layers=3, percep_is_l1=1,
             pool_size=50, DG_ratio=1, lr=2e-4, beta1=0.5, lr_policy='lambda', lr_decay_iters=50, niter=100,
             niter_decay=0, epoch_count=1, continue_train=False, which_epoch='latest', isTrain=True,
             local_rank='cpu', gpu='cpu', distributed=False, opt_level='O0', seed=49, gpu_ids=[],
             checkpoints_dir='./checkpoints', name='oracle')
    o.update(over)
    return types.SimpleNamespace(**o)

