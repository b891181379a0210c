"""On-device replacement for vame/model/dataloader.py.

The reference's ``SEQUENCE_DATASET`` slices one random (F, 2T) window per ``__getitem__`` in Python (index ignored, start
drawn with ``np.random.choice``, dataloader.py:45-56) and the DataLoader collates float64 on the host: ~11 k windows/s at
B=256, i.e. ~23 ms per batch — ten times the fused train step.  ``DeviceWindowSampler`` keeps the z-scored series resident
in HBM and gathers a whole batch of windows with one device-side gather; it yields (B, F, 2T) batches like the
reference loader, so ``vame_b200.rnn_vae.train`` / ``test`` consume it unchanged.
"""
import os

import numpy as np
import torch


class DeviceWindowSampler:
    """Iterable of ``len(series) // batch_size`` random batches (DataLoader(shuffle=True, drop_last=True) semantics with the
    reference dataset's "every item is a fresh random window" behaviour)."""

    def __init__(self, path_to_file, data, train, temporal_window, batch_size, device="cuda", seed=None):
        X = np.load(os.path.join(path_to_file, data))
        if X.shape[0] > X.shape[1]:
            X = X.T                                             # dataloader.py:22-23
        self.data_points = X.shape[1]
        mean_p, std_p = os.path.join(path_to_file, "seq_mean.npy"), os.path.join(path_to_file, "seq_std.npy")
        if train and not os.path.exists(mean_p):                # dataloader.py:27-32
            self.mean, self.std = np.mean(X), np.std(X)
            np.save(mean_p, self.mean)
            np.save(std_p, self.std)
        else:
            self.mean, self.std = np.load(mean_p), np.load(std_p)
        self.temporal_window = int(temporal_window)
        self.batch_size = int(batch_size)
        self.device = torch.device(device)
        z = (X - self.mean) / self.std                          # dataloader.py:54 (applied once instead of per item)
        self.series = torch.from_numpy(np.ascontiguousarray(z.T)).to(self.device, torch.float32)      # (N, F)
        self.gen = torch.Generator(device=self.device)
        if seed is not None:
            self.gen.manual_seed(int(seed))
        else:
            self.gen.seed()
        self._offs = torch.arange(self.temporal_window, device=self.device)

    def __len__(self):
        return self.data_points // self.batch_size              # drop_last=True over len(dataset) = data_points items

    def __iter__(self):
        n_start = self.data_points - self.temporal_window       # np.random.choice(nf - temp_window): starts in [0, n_start)
        for _ in range(len(self)):
            starts = torch.randint(0, n_start, (self.batch_size,), device=self.device, generator=self.gen)
            idx = starts[:, None] + self._offs[None, :]
            yield self.series[idx].permute(0, 2, 1)             # (B, F, 2T) view like the reference loader's batches
