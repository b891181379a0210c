"""Host-side mirror of the inference hot path of vame/analysis/pose_segmentation.py:
``load_model`` (:27-64) and ``embedd_latent_vectors`` (:67-101).

The reference embeds a recording with a batch-size-1 Python loop (one H2D copy, ~10 kernel launches and one D2H copy per
frame).  Here the whole (F, N) series is uploaded once, the layer-0 input projection is computed once per FRAME, and the
N - T stride-1 windows are pushed through the encoder + Lambda mean in large batches by the CUDA library
(vame_embed_windows); windows are independent (h0 = 0), so this is exactly the same arithmetic.
"""
import os

import numpy as np
import torch

from ._lib import VameB200Error
from .rnn_model import RNN_VAE


def load_model(cfg, model_name, fixed):
    """Build the model from the project config, load best_model/<model>_<Project>.pkl, eval mode  [:27-64]."""
    if not torch.cuda.is_available():
        raise VameB200Error("vame_b200.load_model needs a CUDA device (there is no CPU execution path)")
    NUM_FEATURES = cfg['num_features']
    if fixed == False:  # noqa: E712
        NUM_FEATURES = NUM_FEATURES - 2
    model = RNN_VAE(cfg['time_window'] * 2, cfg['zdims'], NUM_FEATURES, cfg['prediction_decoder'], cfg['prediction_steps'],
                    cfg['hidden_size_layer_1'], cfg['hidden_size_layer_2'], cfg['hidden_size_rec'], cfg['hidden_size_pred'],
                    cfg['dropout_encoder'], cfg['dropout_rec'], cfg['dropout_pred'], cfg['softplus']).cuda()
    path = os.path.join(cfg['project_path'], 'model', 'best_model', model_name + '_' + cfg['Project'] + '.pkl')
    model.load_state_dict(torch.load(path, map_location='cuda'))
    model.eval()
    return model


def embed_series(model, data, temp_win, chunk=9472, shard=None):
    """(F, N) array -> (N - temp_win, Z) float32 numpy array of latent means.
    shard = (rank, world): embed only this rank's contiguous slice of the window range (no collective on the data path)."""
    eng = model.engine
    series = torch.from_numpy(np.ascontiguousarray(np.asarray(data).T)).to(torch.float32)
    series = series.pin_memory().to(eng.device, non_blocking=True) if torch.cuda.is_available() else series
    n_win = data.shape[1] - temp_win                      # range(data.shape[1] - temp_win): pose_segmentation.py:87
    first, count = 0, max(n_win, 0)
    if shard is not None:
        rank, world = shard
        per = (count + world - 1) // world
        first = min(count, rank * per)
        count = min(per, count - first)
    mu = eng.embed(series, first_window=first, n_windows=count, chunk=chunk)
    return mu.cpu().numpy()


def embedd_latent_vectors(cfg, files, model, fixed):
    """Same signature / return value as the reference: list of (N - T, zdims) float32 arrays, one per file  [:67-101]."""
    project_path = cfg['project_path']
    temp_win = cfg['time_window']
    num_features = cfg['num_features']
    if fixed == False:  # noqa: E712
        num_features = num_features - 2
    latent_vector_files = []
    was_training = model.training
    model.eval()
    for file in files:
        print('Embedding of latent vector for file %s' % file)
        data = np.load(os.path.join(project_path, 'data', file, file + '-PE-seq-clean.npy'))
        if data.shape[0] != num_features:
            raise ValueError("expected %d features, file %s has %d" % (num_features, file, data.shape[0]))
        with torch.no_grad():
            latent_vector_files.append(embed_series(model, data, temp_win))
    model.train(was_training)
    return latent_vector_files


# ---- k-means parameterization ------------------------------------------------------------------------
def consecutive(data, stepsize=1):
    """pose_segmentation.py:103-105"""
    data = data[:]
    return np.split(data, np.where(np.diff(data) != stepsize)[0] + 1)


def get_motif_usage(label):
    """Motif usage counts with zero-filled gaps between the observed labels  [pose_segmentation.py:108-125]."""
    values, counts = np.unique(label, return_counts=True)
    cons = consecutive(values)
    if len(cons) == 1:
        return counts
    usage = list(counts)
    for i in range(len(cons) - 1):
        gap = (cons[i + 1][0] - cons[i][-1]) - 1
        for j in range(1, gap + 1):
            usage.insert(cons[i][-1] + j, 0)
    return np.array(usage)


def same_parameterization(cfg, files, latent_vector_files, states, parameterization):
    """One k-means over the concatenated latent vectors of all files  [pose_segmentation.py:127-170]; the reference hard-codes
    random_state=42, n_init=20 here (:141).  Only the "kmeans" parameterization is accelerated."""
    if parameterization != "kmeans":
        raise VameB200Error("vame_b200.same_parameterization: only parameterization='kmeans' is implemented (got %r)" % (parameterization,))
    from .kmeans import DeviceKMeans
    latent_vector_cat = np.concatenate(latent_vector_files, axis=0)
    print("Using kmeans as parameterization!")
    km = DeviceKMeans(states, random_state=42, n_init=20).fit(latent_vector_cat)
    clust_center = km.cluster_centers_
    label = km.predict(latent_vector_cat).cpu().numpy()
    labels, cluster_centers, motif_usages = [], [], []
    idx = 0
    for i, _file in enumerate(files):
        file_len = latent_vector_files[i].shape[0]
        labels.append(label[idx:idx + file_len])
        cluster_centers.append(clust_center)
        motif_usages.append(get_motif_usage(label[idx:idx + file_len]))
        idx += file_len
    return labels, cluster_centers, motif_usages


def individual_parameterization(cfg, files, latent_vector_files, cluster):
    """One k-means per file  [pose_segmentation.py:173-196] (the reference reads the seed from the key
    'random_state_kmeans: ', trailing colon and space included; both spellings are accepted here)."""
    from .kmeans import DeviceKMeans
    random_state = cfg["random_state_kmeans: "] if "random_state_kmeans: " in cfg else cfg["random_state_kmeans"]
    n_init = cfg["n_init_kmeans"]
    labels, cluster_centers, motif_usages = [], [], []
    for i, file in enumerate(files):
        print(file)
        km = DeviceKMeans(cluster, random_state=random_state, n_init=n_init).fit(latent_vector_files[i])
        label = km.predict(latent_vector_files[i]).cpu().numpy()
        motif_usages.append(get_motif_usage(label))
        labels.append(label)
        cluster_centers.append(km.cluster_centers_)
    return labels, cluster_centers, motif_usages
