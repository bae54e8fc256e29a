"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's retrieval scoring and feature layout.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module. It follows
/root/reference/evaluate.py:
  * :78-86  features of one tolerance = the per-batch model outputs [T, n, 64] concatenated along axis -2; the
            tolerances are stacked in front -> features[tol, T, N, 64] (:88-94 writes exactly this array to HDF5);
  * :326    `features /= np.linalg.norm(features, axis=-2, keepdims=True) + 1e-7` - the norm runs over the SAMPLE axis;
  * :339    `scores = queries.dot(db.T)` with queries = db = features[..] of one (tol, t1) plane (:328, :348-355).
The reference has no test or golden vector for this path and needs h5py (absent here); it is pinned by these three
lines, which are plain numpy calls, i.e. the restatement IS the reference's arithmetic.
"""
import numpy as np


def stack_features(per_tol_batches):
    """per_tol_batches[tol][batch] = array [T, n_b, D]  ->  features [tol, T, N, D] (evaluate.py:78-86)."""
    return np.stack([np.concatenate(batches, -2) for batches in per_tol_batches])


def normalize(features):
    """evaluate.py:326 (out of place)."""
    features = np.array(features, dtype=np.float32, copy=True)
    features /= np.linalg.norm(features, axis=-2, keepdims=True) + 1e-7
    return features


def scores(queries, db):
    """evaluate.py:339."""
    return queries.dot(db.T)
