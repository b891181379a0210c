"""CPU checks of the host-side mirror of the reference interface (names, init, state_dict, install hooks)."""
import sys

import numpy as np
import pytest
import torch

from oracle import vame_oracle as vo
from oracle.ref_shim import reference_available


def test_state_dict_keys_shapes_and_init_match_port():
    from vame_b200.engine import Engine, state_dict_names
    from vame_b200.rnn_model import RNN_VAE
    for fut in (True, False):
        torch.manual_seed(19)
        m = RNN_VAE(60, 30, 24, fut, 15, 256, 256, 256, 256, 0, 0, 0, False)
        torch.manual_seed(19)
        port = vo.RefPort(60, 30, 24, fut, 15, hidden=256)
        sd, psd = m.state_dict(), port.state_dict()
        assert list(sd.keys()) == list(psd.keys()) == state_dict_names(fut)
        for k in sd:
            assert torch.equal(sd[k], psd[k]), k
        e = Engine(24, 30, 30, 256, 256, 256, fut, 15)
        assert [tuple(sd[k].shape) for k in sd] == [tuple(s) for s in e.shapes]
        assert m.seq_len == 30 and m.FUTURE_DECODER == fut
        assert hasattr(m, "decoder_future") == fut


def test_kl_annealing_matches_reference_semantics():
    from vame_b200.rnn_vae import kl_annealing
    assert kl_annealing(1, 2, 4, "linear") == 0
    assert kl_annealing(3, 2, 4, "linear") == 0.25
    assert kl_annealing(10, 2, 4, "linear") == 1
    assert abs(kl_annealing(5, 2, 4, "sigmoid") - 1 / (1 + np.exp(-0.9))) < 1e-12
    with pytest.raises(NotImplementedError):
        kl_annealing(5, 2, 4, "cosine")


def test_generic_losses_match_port_on_cpu():
    from vame_b200 import rnn_vae as rv
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(4, 5, 3, generator=g), torch.randn(4, 5, 3, generator=g)
    for red in ("sum", "mean"):
        assert torch.allclose(rv.reconstruction_loss(a, b, red), vo.reconstruction_loss(a, b, red))
    mu, lv = torch.randn(6, 4, generator=g), torch.randn(6, 4, generator=g)
    assert torch.allclose(rv.kullback_leibler_loss(mu, lv), vo.kullback_leibler_loss(mu, lv))


@pytest.mark.reference
@pytest.mark.skipif(not reference_available(), reason="reference not mounted")
def test_install_rebinds_hot_path_names():
    from oracle.ref_shim import load_reference
    load_reference()
    from vame_b200 import install, rnn_model, rnn_vae, pose_segmentation
    saved = {}
    mods = ["vame.model.rnn_model", "vame.model.rnn_vae", "vame.analysis.pose_segmentation", "vame.model.evaluate",
            "vame.analysis.generative_functions", "vame.model.create_training"]
    for m in mods:
        saved[m] = dict(sys.modules[m].__dict__)
    try:
        done = install.install()
        assert ("vame.model.rnn_vae", "train") in done and ("vame.analysis.pose_segmentation", "embedd_latent_vectors") in done
        assert sys.modules["vame.model.rnn_vae"].RNN_VAE is rnn_model.RNN_VAE
        assert sys.modules["vame.model.rnn_vae"].train is rnn_vae.train
        assert sys.modules["vame.model.rnn_vae"].cluster_loss is rnn_vae.cluster_loss
        from vame_b200 import dataloader
        assert sys.modules["vame.model.rnn_vae"].SEQUENCE_DATASET is dataloader.SEQUENCE_DATASET      # the loader seam
        assert sys.modules["vame.model.rnn_vae"].Data.DataLoader is dataloader.Data.DataLoader
        assert sys.modules["vame.model.rnn_vae"].Data.Dataset is torch.utils.data.Dataset               # everything else forwards
        assert sys.modules["vame.analysis.pose_segmentation"].RNN_VAE is rnn_model.RNN_VAE
        assert sys.modules["vame.analysis.pose_segmentation"].load_model is pose_segmentation.load_model
        assert sys.modules["vame.model.evaluate"].RNN_VAE is rnn_model.RNN_VAE
        ct = sys.modules["vame.model.create_training"]
        assert ("vame.model.create_training", "traindata_fixed") in done and ct.traindata_fixed is not ct._ref_traindata_fixed
        assert ct.create_trainset.__module__ == "vame.model.create_training"       # the driver function stays the reference's
    finally:
        for m in mods:
            sys.modules[m].__dict__.clear()
            sys.modules[m].__dict__.update(saved[m])


@pytest.mark.reference
@pytest.mark.skipif(not reference_available(), reason="reference not mounted")
def test_sequence_dataset_mirror_matches_reference_class(tmp_path):
    """vame_b200.dataloader.SEQUENCE_DATASET vs the reference class on the same files: statistics files, length, item
    arithmetic (same numpy RNG state -> same window), and the plain-DataLoader fallback of the Data shim on a CPU box."""
    import os
    from oracle.ref_shim import load_reference
    load_reference()
    ref_dl = sys.modules["vame.model.dataloader"]
    from vame_b200 import dataloader as dl
    rng = np.random.default_rng(3)
    X = np.cumsum(rng.standard_normal((5000, 8)), axis=0)            # (N, F): both classes transpose it (dataloader.py:22-23)
    da, db = str(tmp_path / "a") + os.sep, str(tmp_path / "b") + os.sep
    os.makedirs(da), os.makedirs(db)
    for d in (da, db):
        np.save(d + "train_seq.npy", X)
    ref = ref_dl.SEQUENCE_DATASET(da, data="train_seq.npy", train=True, temporal_window=40)
    ours = dl.SEQUENCE_DATASET(db, data="train_seq.npy", train=True, temporal_window=40)
    assert len(ref) == len(ours) == 5000
    assert np.load(da + "seq_mean.npy") == np.load(db + "seq_mean.npy") and np.load(da + "seq_std.npy") == np.load(db + "seq_std.npy")
    np.random.seed(11)
    a = ref[0]
    np.random.seed(11)
    b = ours[0]
    assert a.dtype == b.dtype == torch.float64 and torch.equal(a, b)
    if not torch.cuda.is_available():
        loader = dl.Data.DataLoader(ours, batch_size=16, shuffle=True, drop_last=True)
        assert isinstance(loader, torch.utils.data.DataLoader) and tuple(next(iter(loader)).shape) == (16, 8, 40)
