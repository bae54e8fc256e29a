"""feature layout features[tol, T, N, 64] and retrieval scoring (reference evaluate.py:78-94, :326, :339).
CPU: the numpy oracle on hand-checkable inputs. GPU: csrc/retrieval.cu and node_b200.retrieval against the oracle."""
import numpy as np
import pytest
import torch

from oracle import retrieval_port


def test_oracle_normalises_over_the_sample_axis():
    f = np.zeros((1, 2, 3, 2), dtype=np.float32)            # [tol, T, N, D]
    f[0, 0, :, 0] = [3.0, 4.0, 0.0]
    f[0, 0, :, 1] = [1.0, 0.0, 0.0]
    f[0, 1, :, 0] = [0.0, 0.0, 2.0]
    out = retrieval_port.normalize(f)
    np.testing.assert_allclose(out[0, 0, :, 0], [0.6, 0.8, 0.0], rtol=1e-6)        # norm 5 over the three samples
    np.testing.assert_allclose(out[0, 0, :, 1], [1.0, 0.0, 0.0], rtol=1e-6)
    np.testing.assert_allclose(out[0, 1, :, 0], [0.0, 0.0, 1.0], rtol=1e-6)
    assert np.all(out[0, 1, :, 1] == 0.0)                                          # 0 / (0 + 1e-7)
    s = retrieval_port.scores(out[0, 0], out[0, 0])
    np.testing.assert_allclose(s, [[1.36, 0.48, 0.0], [0.48, 0.64, 0.0], [0.0, 0.0, 0.0]], rtol=1e-6)


def test_oracle_feature_layout():
    rng = np.random.default_rng(0)
    batches = [[rng.standard_normal((4, n, 8)).astype(np.float32) for n in (5, 5, 3)] for _ in range(2)]
    f = retrieval_port.stack_features(batches)
    assert f.shape == (2, 4, 13, 8)
    assert np.array_equal(f[1, 2, 5:10], batches[1][1][2])


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(2, 3, 1000, 64), (1, 1, 130, 64), (10, 257, 10), (1, 2, 77, 24)])
def test_feature_normalize_matches_oracle(native_lib, shape):
    from node_b200 import retrieval
    rng = np.random.default_rng(3)
    f = (rng.standard_normal(shape) * rng.uniform(0.1, 5.0, size=shape[-1])).astype(np.float32)
    f[..., 1] = 0.0                                          # a dead feature dimension: 0 / 1e-7
    ref = retrieval_port.normalize(f)
    got, norms = retrieval.normalize_features(torch.from_numpy(f).cuda())
    torch.cuda.synchronize()
    np.testing.assert_allclose(norms.cpu().numpy(), np.linalg.norm(f.astype(np.float64), axis=-2, keepdims=True), rtol=1e-6, atol=0)
    err = np.abs(got.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err <= 1e-6, err


@pytest.mark.gpu
@pytest.mark.parametrize('nq,ns,d', [(1000, 1000, 64), (130, 257, 64), (1, 5, 64), (300, 129, 10), (64, 64, 7)])
def test_retrieval_scores_match_oracle(native_lib, nq, ns, d):
    from node_b200 import retrieval
    rng = np.random.default_rng(5)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    db = rng.standard_normal((ns, d)).astype(np.float32)
    ref = retrieval_port.scores(q.astype(np.float64), db.astype(np.float64))
    ref32 = retrieval_port.scores(q, db)
    got = retrieval.retrieval_scores(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda()).cpu().numpy()
    assert got.shape == (nq, ns)
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 2e-6 * scale
    assert np.abs(got - ref).max() <= 2.0 * np.abs(ref32 - ref).max() + 1e-7 * scale      # as close to exact as numpy's own fp32 product


@pytest.mark.gpu
def test_feature_store_matches_reference_layout(native_lib):
    """extract_features fills features[tol, T, N, 64] in place; the reference concatenates the per-batch model outputs
    (evaluate.py:78-86). Retrieval on top of it: identical ranking of every query's nearest samples."""
    from node_b200 import models, retrieval
    torch.manual_seed(0)
    net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3).eval().cuda()
    net.to_features_extractor()
    batches = [torch.rand(n, 3, 32, 32, device='cuda') for n in (16, 16, 9)]
    tols, t1s = [1e-3, 1e-2], np.linspace(0, 1, 4).tolist()
    store = retrieval.extract_features(net, batches, tols, t1s)
    per_tol = []
    with torch.no_grad():
        for tol in tols:
            net.odeblock.tol = tol
            per_tol.append([net(x).cpu().numpy() for x in batches])
    ref = retrieval_port.stack_features(per_tol)
    got = store.to_host()
    assert got.shape == ref.shape == (2, 4, 41, 64)
    assert np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max()
    nref = retrieval_port.normalize(ref)
    ngot, _ = retrieval.normalize_features(store.features)
    s_ref = retrieval_port.scores(nref[0, -1], nref[0, -1])
    s_got = retrieval.retrieval_scores(ngot[0, -1], ngot[0, -1]).cpu().numpy()
    assert np.abs(s_got - s_ref).max() <= 1e-5 * np.abs(s_ref).max()
