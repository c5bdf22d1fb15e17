"""History buffer of generated images (reference: util/image_pool.py:7-34).

Host-side bookkeeping only: the stored items are per-image device tensors, the draws come from Python's
``random`` exactly like the reference so that a seeded run replays the same swaps."""
import random

import torch


class ImagePool():
    def __init__(self, pool_size):
        self.pool_size = pool_size
        if self.pool_size > 0:
            self.num_imgs = 0
            self.images = []

    def query(self, images):
        if self.pool_size == 0:
            return images
        picked = []
        for image in images:
            image = torch.unsqueeze(image, 0)
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(image)
                picked.append(image)
            elif random.uniform(0, 1) > 0.5:
                slot = random.randint(0, self.pool_size - 1)
                old = self.images[slot].clone()
                self.images[slot] = image
                picked.append(old)
            else:
                picked.append(image)
        return torch.cat(picked, 0)
