"""Drop-in for the reference's ``models/base_model.py`` (:8-87): option plumbing, checkpoint save / load with the
reference's file names (``<label>_net_<name>.pth`` holding the un-wrapped module's fp32 ``state_dict``) and the
per-epoch learning-rate step. No APEX: there is no ``*_net_amp.pth`` to write (bf16 needs no loss scaling) and a
stray one on disk is ignored on load."""
import os

import torch
import torch.nn as nn


class BaseModel(nn.Module):
    def __init__(self, opt):
        super(BaseModel, self).__init__()
        self.opt = opt
        self.gpu_ids = [opt.local_rank]
        self.isTrain = opt.isTrain
        self.Tensor = torch.cuda.FloatTensor if self.gpu_ids else torch.Tensor
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.master = opt.local_rank == 0

    def name(self):
        return 'BaseModel'

    def set_input(self, input):
        self.input = input

    def forward(self):
        pass

    def test(self):
        pass

    def get_image_paths(self):
        pass

    def optimize_parameters(self):
        pass

    def get_current_visuals(self):
        return self.input

    def get_current_errors(self):
        return {}

    def save(self, label):
        pass

    def save_network(self, network, network_label, epoch_label, gpu_ids):
        if self.master:
            os.makedirs(self.save_dir, exist_ok=True)
            save_path = os.path.join(self.save_dir, '%s_net_%s.pth' % (epoch_label, network_label))
            net = network.module if hasattr(network, 'module') else network
            torch.save({k: v.detach().cpu() for k, v in net.state_dict().items()}, save_path)

    def load_network(self):
        opt = self.opt
        root = os.path.join('checkpoints', opt.name)
        for fname in os.listdir(root):
            if opt.which_epoch not in fname or 'amp' in fname:
                continue
            name = fname[0:-4].replace("%s_net_" % opt.which_epoch, '')
            sub_model = getattr(self, name)
            sub_model.load_state_dict(torch.load(os.path.join(root, fname), map_location='cpu'))
            print("loading weights for %s" % name)

    def update_learning_rate(self):
        for scheduler in self.schedulers:
            scheduler.step()
        lr = self.optimizers[0].param_groups[0]['lr']
        print('learning rate = %.7f' % lr)
