"""History buffer of generated images (reference: util/image_pool.py:7-34).

Host-side bookkeeping only: the draws come from Python's ``random`` exactly like the reference so that a seeded
run replays the same swaps. The stored images live in one preallocated ``[pool_size, C, H, W]`` device buffer (the
reference keeps views into every batch it ever saw, which pins whole batches in memory and churns the allocator);
``query(images, out=...)`` writes the picked batch into a caller-owned buffer."""
import random

import torch


class ImagePool():
    def __init__(self, pool_size):
        self.pool_size = pool_size
        if self.pool_size > 0:
            self.num_imgs = 0
            self.store = None

    def query(self, images, out=None):
        if self.pool_size == 0:
            if out is not None:
                out.copy_(images)
                return out
            return images
        if out is None:
            out = torch.empty_like(images)
        if getattr(self, 'store', None) is None or self.store.shape[1:] != images.shape[1:] or \
                self.store.device != images.device:
            self.store = torch.empty((self.pool_size,) + tuple(images.shape[1:]), dtype=images.dtype,
                                     device=images.device)
            self.num_imgs = 0
        for i in range(images.shape[0]):
            image = images[i]
            if self.num_imgs < self.pool_size:
                self.store[self.num_imgs].copy_(image)
                self.num_imgs += 1
                out[i].copy_(image)
            elif random.uniform(0, 1) > 0.5:
                slot = random.randint(0, self.pool_size - 1)
                out[i].copy_(self.store[slot])
                self.store[slot].copy_(image)
            else:
                out[i].copy_(image)
        return out
