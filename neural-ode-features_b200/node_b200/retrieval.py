"""Edges after the feature extractor: the reference's HDF5 feature layout fed from device buffers and
its retrieval scoring as device kernels (csrc/retrieval.cu).

Mirrors /root/reference/evaluate.py:
  * `features()` (:24-94): for every tolerance, the feature extractor's outputs [T, n, 64] of all batches are concatenated
    along the sample axis and the tolerances stacked -> `features[tol, T, N, 64]`, written with `y_true`, `tols`, `t1s`.
    Here the classifier head writes every batch straight into its slice of ONE device buffer of that layout
    (FeatureStore); the host sees it through a single device -> pinned-host copy.
  * `retrieval()` (:326, :339): `features / (norm over the sample axis + 1e-7)` and `scores = queries . db^T`.
"""
import numpy as np
import torch

from . import caller_ops, native


class FeatureStore(object):
    """Device buffer `features[tol, T, N, D]` (evaluate.py:88-94) that the feature extractor fills in place."""

    def __init__(self, n_tol, n_times, n_samples, dim=64, device='cuda'):
        self.features = torch.zeros((n_tol, n_times, n_samples, dim), dtype=torch.float32, device=device)
        self.filled = [0] * n_tol

    def slot(self, tol_index, t_index, n):
        """The contiguous [n, D] slice the next batch of tolerance `tol_index` occupies at output time `t_index`."""
        off = self.filled[tol_index]
        return self.features[tol_index, t_index, off:off + n]

    def advance(self, tol_index, n):
        self.filled[tol_index] += n

    def to_host(self):
        """One device -> pinned host copy of the whole array (numpy view of the pinned buffer)."""
        host = torch.empty(self.features.shape, dtype=torch.float32, pin_memory=self.features.is_cuda)
        host.copy_(self.features)
        return host.numpy()

    def save(self, path, y_true, tols, t1s):
        """The reference's datasets `features`, `y_true`, `tols`, `t1s` (evaluate.py:90-94): HDF5 when h5py is importable,
        else an .npz with the same keys."""
        arrays = dict(features=self.to_host(), y_true=np.asarray(y_true), tols=np.asarray(tols), t1s=np.asarray(t1s))
        try:
            import h5py
        except ImportError:
            np.savez(path if path.endswith('.npz') else path + '.npz', **arrays)
            return
        with h5py.File(path, 'w') as f:
            for k, v in arrays.items():
                f[k] = v


def extract_features(model, batches, tols, t1s):
    """evaluate.py:62-86 for an ODENet turned into a feature extractor (`to_features_extractor()`): for every tolerance run
    every batch and store the pooled features of every output time. `batches` is a sequence of input tensors (already on
    the model's device). Returns the FeatureStore."""
    dev = next(model.parameters()).device
    n_total = sum(int(b.shape[0]) for b in batches)
    model.odeblock.t1 = list(t1s)
    n_times = int(model.odeblock.integration_time.numel())
    dim = model.odeblock.odefunc.norm1.num_channels
    store = FeatureStore(len(tols), n_times, n_total, dim, dev)
    with torch.no_grad():
        for ti, tol in enumerate(tols):
            model.odeblock.tol = tol
            for x in batches:
                ys = model.odeblock(model.downsample(x))          # [T, n, C, H, W]
                n = int(x.shape[0])
                for k in range(n_times):
                    caller_ops.head(model.classifier.module, ys[k], out=store.slot(ti, k, n))
                store.advance(ti, n)
    return store


@native.on_device_of(0)
def normalize_features(features, out=None):
    """evaluate.py:326 on a device tensor [..., N, D]: features / (norm over axis -2 + 1e-7). Returns (normalised, norms)."""
    if not (features.is_cuda and features.dtype == torch.float32 and features.dim() >= 2):
        raise TypeError('normalize_features: a CUDA fp32 tensor [..., N, D] is required (no CPU fallback)')
    f = features.contiguous()
    N, D = int(f.shape[-2]), int(f.shape[-1])
    planes = f.numel() // (N * D)
    out = torch.empty_like(f) if out is None else out
    norms = torch.empty(f.shape[:-2] + (1, D), dtype=torch.float32, device=f.device)
    native.check(native.lib().node_b200_feature_normalize(native.ptr(f), native.ptr(out), native.ptr(norms), planes, N, D,
                                                          native.stream_ptr()), 'feature_normalize')
    return out, norms


@native.on_device_of(0)
def retrieval_scores(queries, db):
    """evaluate.py:339: scores[q, s] = <queries[q], db[s]> for device tensors [nq, D], [ns, D] -> [nq, ns]."""
    for t in (queries, db):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
            raise TypeError('retrieval_scores: CUDA fp32 matrices are required (no CPU fallback)')
    if queries.shape[1] != db.shape[1]:
        raise ValueError('retrieval_scores: feature dimensions differ')
    q, d = queries.contiguous(), db.contiguous()
    out = torch.empty((q.shape[0], d.shape[0]), dtype=torch.float32, device=q.device)
    native.check(native.lib().node_b200_retrieval_scores(native.ptr(q), native.ptr(d), native.ptr(out), int(q.shape[0]), int(d.shape[0]),
                                                         int(q.shape[1]), native.stream_ptr()), 'retrieval_scores')
    return out
