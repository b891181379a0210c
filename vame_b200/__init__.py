"""vame_b200 — B200-native (sm_100a) RNN-VAE hot path behind VAME's Python surface.

Public surface mirrors the reference for the hot path only:
  vame_b200.rnn_model          Encoder / Lambda / Decoder / Decoder_Future / RNN_VAE
  vame_b200.rnn_vae            loss functions, kl_annealing, train, test
  vame_b200.pose_segmentation  load_model, embedd_latent_vectors
  vame_b200.install.install()  rebinds those names inside an imported, unmodified ``vame`` package
The arithmetic lives in libvame_b200.so (C-ABI in include/vame_b200.h); importing this package does not load it,
the first use does, and it raises if the library is missing (no CPU / PyTorch fallback).
"""
__version__ = "0.1.0"
