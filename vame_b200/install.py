"""Drop the B200 hot path in behind an UNMODIFIED VAME installation.

``install()`` rebinds, in the already-imported reference modules, exactly the names on the hot path
(SURVEY.md §8b): the five model classes, the four loss functions, ``train`` / ``test``, the dataset / loader pair that
``train_model`` builds (``SEQUENCE_DATASET``, ``Data.DataLoader`` -> on-device window sampler) and
``load_model`` / ``embedd_latent_vectors``, the k-means parameterization and the arithmetic of ``create_trainset``
(``traindata_fixed`` / ``traindata_aligned``).  Everything else (config handling, csv / alignment preprocessing, train_model's epoch
loop, checkpoint files, clustering, plotting) keeps running the reference's own code, so existing projects and
``.pkl`` checkpoints work unchanged.
"""
import sys

from . import pose_segmentation as _ps
from . import rnn_model as _rm
from . import rnn_vae as _rv

MODEL_NAMES = ("Encoder", "Lambda", "Decoder", "Decoder_Future", "RNN_VAE")
TRAIN_NAMES = ("reconstruction_loss", "future_reconstruction_loss", "cluster_loss", "kullback_leibler_loss",
               "kl_annealing", "gaussian", "train", "test")
CONSUMERS = ("vame.model.rnn_vae", "vame.analysis.pose_segmentation", "vame.model.evaluate",
             "vame.analysis.generative_functions", "vame.analysis.segment_behavior")


def install(verbose=False):
    """Patch the imported ``vame`` package in place.  Returns the list of (module, name) pairs that were rebound."""
    if "vame" not in sys.modules:
        import vame  # noqa: F401  (must be importable; the reference package itself is not modified)
    done = []
    ref_rm = sys.modules.get("vame.model.rnn_model")
    if ref_rm is not None:
        for n in MODEL_NAMES:
            setattr(ref_rm, n, getattr(_rm, n))
            done.append((ref_rm.__name__, n))
    for modname in CONSUMERS:
        mod = sys.modules.get(modname)      # NB: vame.analysis.pose_segmentation the ATTRIBUTE is a function; use sys.modules
        if mod is None:
            continue
        for n in MODEL_NAMES:
            if hasattr(mod, n):
                setattr(mod, n, getattr(_rm, n))
                done.append((modname, n))
    rv = sys.modules.get("vame.model.rnn_vae")
    if rv is not None:
        for n in TRAIN_NAMES:
            setattr(rv, n, getattr(_rv, n))
            done.append((rv.__name__, n))
        rv.use_gpu = True
        # the loader seam (rnn_vae.py:23,326-330): train_model builds SEQUENCE_DATASET(...) and Data.DataLoader(...) from these
        # module globals; ours keep the series in HBM and sample every batch with one kernel inside the captured step
        from . import dataloader as _dl
        rv.SEQUENCE_DATASET = _dl.SEQUENCE_DATASET
        rv.Data = _dl.Data
        done += [(rv.__name__, "SEQUENCE_DATASET"), (rv.__name__, "Data.DataLoader")]
    ps = sys.modules.get("vame.analysis.pose_segmentation")
    if ps is not None:
        for n in ("load_model", "embedd_latent_vectors", "individual_parameterization"):
            setattr(ps, n, getattr(_ps, n))
            done.append((ps.__name__, n))
        # k-means parameterization on the device; the HMM branch of same_parameterization stays the reference's
        if not hasattr(ps, "_ref_same_parameterization"):
            ps._ref_same_parameterization = ps.same_parameterization

        def same_parameterization(cfg, files, latent_vector_files, states, parameterization):
            if parameterization == "kmeans":
                return _ps.same_parameterization(cfg, files, latent_vector_files, states, parameterization)
            return ps._ref_same_parameterization(cfg, files, latent_vector_files, states, parameterization)

        ps.same_parameterization = same_parameterization
        done.append((ps.__name__, "same_parameterization"))
    # training-set preparation (create_trainset step): the arithmetic runs on the device, files and formats are the reference's
    ct = sys.modules.get("vame.model.create_training")
    if ct is not None:
        from . import create_training as _ct
        if not hasattr(ct, "_ref_traindata_fixed"):
            ct._ref_traindata_fixed, ct._ref_traindata_aligned = ct.traindata_fixed, ct.traindata_aligned

        def _wrap(ours, ref):
            def f(cfg, files, testfraction, num_features, savgol_filter, check_parameter):
                if check_parameter:                      # the matplotlib inspection plots stay the reference's
                    return ref(cfg, files, testfraction, num_features, savgol_filter, check_parameter)
                return ours(cfg, files, testfraction, num_features, savgol_filter, check_parameter)
            return f
        ct.traindata_fixed = _wrap(_ct.traindata_fixed, ct._ref_traindata_fixed)
        ct.traindata_aligned = _wrap(_ct.traindata_aligned, ct._ref_traindata_aligned)
        done += [(ct.__name__, "traindata_fixed"), (ct.__name__, "traindata_aligned")]
    if verbose:
        for m, n in done:
            print("vame_b200: %s.%s -> B200 path" % (m, n))
    return done
