"""On-device replacement for vame/model/dataloader.py + the torch DataLoader that train_model builds from it
(vame/model/rnn_vae.py:326-330).

The reference's ``SEQUENCE_DATASET`` slices ONE random (F, 2T) window per ``__getitem__`` in Python (the index is ignored, the
start is drawn with ``np.random.choice``, dataloader.py:45-56) and the DataLoader collates float64 on the host: ~11 k windows/s
at B = 256, i.e. ~23 ms per batch - 24 times the fused train step.  Here

* ``SEQUENCE_DATASET`` mirrors the reference class (constructor arguments, ``seq_mean.npy`` / ``seq_std.npy`` files, printed
  messages, ``__len__``, ``__getitem__`` for foreign callers) and keeps the series in the reference's (F, N) float64 layout;
* ``DeviceWindowSampler`` is what ``Data.DataLoader(dataset, batch_size, shuffle=True, drop_last=True)`` returns after
  ``vame_b200.install()``: the series is uploaded once, and every batch is ONE launch of ``vame_sample_windows``
  (csrc/sampler.cu: Philox starts, gather, float64 z-score, float32 cast, data / future split, reparameterisation noise);
  ``vame_b200.rnn_vae.train`` captures that launch inside the CUDA graph of the step, so an epoch is one graph replay per
  batch and no host work per window at all.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib as L


class SEQUENCE_DATASET(torch.utils.data.Dataset):
    """Same constructor and files as vame/model/dataloader.py:18-41."""

    def __init__(self, path_to_file, data, train, temporal_window):
        self.temporal_window = temporal_window
        self.X = np.load(os.path.join(path_to_file, data))
        if self.X.shape[0] > self.X.shape[1]:
            self.X = self.X.T                                   # dataloader.py:22-23
        self.data_points = len(self.X[0, :])
        mean_p, std_p = os.path.join(path_to_file, "seq_mean.npy"), os.path.join(path_to_file, "seq_std.npy")
        if train and not os.path.exists(mean_p):                # dataloader.py:27-32
            print("Compute mean and std for temporal dataset.")
            self.mean = np.mean(self.X)
            self.std = np.std(self.X)
            np.save(mean_p, self.mean)
            np.save(std_p, self.std)
        else:
            self.mean = np.load(mean_p)
            self.std = np.load(std_p)
        if train:
            print('Initialize train data. Datapoints %d' % self.data_points)
        else:
            print('Initialize test data. Datapoints %d' % self.data_points)
        self._dev = {}

    def __len__(self):
        return self.data_points

    def __getitem__(self, index):
        """dataloader.py:45-56 (host path, kept for foreign callers; the train loop never calls it)."""
        start = np.random.choice(self.data_points - self.temporal_window)
        sequence = self.X[:, start:start + self.temporal_window]
        return torch.from_numpy((sequence - self.mean) / self.std)

    def device_series(self, device):
        """(F, N) float64 copy of the series in HBM (uploaded once per device)."""
        device = torch.device(device)
        t = self._dev.get(device)
        if t is None:
            t = torch.from_numpy(np.ascontiguousarray(self.X, dtype=np.float64)).to(device)
            self._dev[device] = t
        return t


class DeviceWindowSampler:
    """Iterable of ``len(dataset) // batch_size`` random batches: DataLoader(shuffle=True, drop_last=True) semantics on top of
    the reference dataset's "every item is a fresh random window" behaviour.  Iterating yields (B, F, 2T) float32 CUDA
    tensors (views of (B, 2T, F) buffers) like the reference loader's batches; ``fill`` writes a batch straight into the
    static buffers of a captured train step."""

    def __init__(self, dataset, batch_size, device="cuda", seed=None):
        if not torch.cuda.is_available():
            raise L.VameB200Error("vame_b200.DeviceWindowSampler needs a CUDA device (there is no CPU fallback)")
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.device = torch.device(device if torch.device(device).index is not None else "cuda:%d" % torch.cuda.current_device())
        self.temporal_window = int(dataset.temporal_window)
        self.data_points = int(dataset.data_points)
        if self.data_points <= self.temporal_window:
            raise ValueError("the series (%d frames) is shorter than one window (%d)" % (self.data_points, self.temporal_window))
        self.series = dataset.device_series(self.device)
        self.num_features = int(self.series.shape[0])
        self.mean, self.std = float(dataset.mean), float(dataset.std)
        self.seed = int(seed) if seed is not None else int.from_bytes(os.urandom(8), "little")
        self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)      # draws so far (advanced on the device)
        self.lib = L.lib()

    @classmethod
    def from_files(cls, path_to_file, data, train, temporal_window, batch_size, device="cuda", seed=None):
        return cls(SEQUENCE_DATASET(path_to_file, data, train, temporal_window), batch_size, device=device, seed=seed)

    def __len__(self):
        return self.data_points // self.batch_size              # drop_last=True over len(dataset) = data_points items

    def fill(self, x, fut=None, eps=None, starts=None, starts_out=None):
        """One batch into x (B, t_data, F) [, fut (B, t_future, F)] [, eps (B, Z)] on the current stream (graph-capturable)."""
        B, t_data = int(x.shape[0]), int(x.shape[1])
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and int(x.shape[2]) == self.num_features
        t_fut = 0
        if fut is not None:
            assert fut.is_contiguous() and fut.dtype == torch.float32 and fut.shape[0] == B and int(fut.shape[2]) == self.num_features
            t_fut = int(fut.shape[1])
        zd = 0
        if eps is not None:
            assert eps.is_contiguous() and eps.dtype == torch.float32 and eps.shape[0] == B
            zd = int(eps.shape[1])
        with torch.cuda.device(self.device):
            L.check(self.lib.vame_sample_windows(L.ptr(self.series), self.data_points, self.num_features, self.temporal_window,
                                                 self.mean, self.std, B, t_data, t_fut, zd, L.ptr(starts),
                                                 ctypes.c_ulonglong(self.seed & 0xFFFFFFFFFFFFFFFF), L.ptr(self.counter), L.ptr(x),
                                                 L.ptr(fut), L.ptr(eps), L.ptr(starts_out), L.cur_stream()), "vame_sample_windows")

    def __iter__(self):
        for _ in range(len(self)):
            buf = torch.empty(self.batch_size, self.temporal_window, self.num_features, device=self.device)
            self.fill(buf)
            yield buf.permute(0, 2, 1)                          # (B, F, 2T) like the reference loader


class _DataNamespace:
    """Stand-in for the ``torch.utils.data`` module object that vame/model/rnn_vae.py binds as ``Data`` (:14):
    ``Data.DataLoader(<our SEQUENCE_DATASET>, batch_size=..., shuffle=True, drop_last=True)`` (rnn_vae.py:329-330) returns the
    device sampler; every other use is forwarded to torch.utils.data."""

    def __getattr__(self, name):
        import torch.utils.data as tud
        return getattr(tud, name)

    @staticmethod
    def DataLoader(dataset, batch_size=1, shuffle=False, drop_last=False, **kw):
        import torch.utils.data as tud
        if isinstance(dataset, SEQUENCE_DATASET) and torch.cuda.is_available() and shuffle and drop_last and not kw:
            return DeviceWindowSampler(dataset, batch_size)
        return tud.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, drop_last=drop_last, **kw)


Data = _DataNamespace()
