"""Property tests of the oracle (CPU, hypothesis): the size-independent laws the GPU tests rely on at full size --
permutation invariance over pixels and samples, linearity in the features, shard-and-sum == whole (the all-reduce
decomposition), adjoint identity of the pooling backward, and exactness of hard-label counts."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import clr_oracle as O

shapes = st.tuples(st.integers(1, 4), st.integers(1, 7), st.integers(1, 6), st.integers(1, 6), st.integers(1, 3))


def _case(shape, seed, hard=True):
    B, C, H, W, K = shape
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, C, H, W))
    y = rng.random((B, K, H, W))
    if hard:
        y = (y > 0.4).astype(np.float64)
    return x, y


@settings(max_examples=40, deadline=None, derandomize=True)
@given(shapes, st.integers(0, 2 ** 31 - 1))
def test_pool_sums_permutation_invariant_and_linear(shape, seed):
    x, y = _case(shape, seed, hard=False)
    B, C, H, W = x.shape
    w = O.weights_complement(y)
    S, N = O.pool_sums(x, w)
    rng = np.random.default_rng(seed + 1)
    perm = rng.permutation(H * W)
    xp = x.reshape(B, C, -1)[:, :, perm].reshape(x.shape)
    wp = w.reshape(B, w.shape[1], -1)[:, :, perm].reshape(w.shape)
    S2, N2 = O.pool_sums(xp, wp)
    assert np.allclose(S, S2, rtol=1e-12, atol=1e-12) and np.allclose(N, N2, rtol=1e-12)
    pb = rng.permutation(B)
    S3, N3 = O.pool_sums(x[pb], w[pb])
    assert np.allclose(S, S3, rtol=1e-12, atol=1e-12) and np.allclose(N, N3, rtol=1e-12)
    S4, _ = O.pool_sums(2.5 * x + 1.0, w)       # linear + the constant picks up the weight sums
    assert np.allclose(S4, 2.5 * S + N[:, None], rtol=1e-11, atol=1e-11)


@settings(max_examples=40, deadline=None, derandomize=True)
@given(shapes, st.integers(0, 2 ** 31 - 1), st.integers(1, 3))
def test_shard_and_sum_equals_whole_batch(shape, seed, nshard):
    x, y = _case(shape, seed)
    w = O.weights_complement(y)
    S, N = O.pool_sums(x, w)
    cuts = np.linspace(0, x.shape[0], nshard + 1).astype(int)
    Ss = sum(O.pool_sums(x[a:b], w[a:b])[0] for a, b in zip(cuts[:-1], cuts[1:]) if b > a)
    Ns = sum(O.pool_sums(x[a:b], w[a:b])[1] for a, b in zip(cuts[:-1], cuts[1:]) if b > a)
    assert np.allclose(S, Ss, rtol=1e-12, atol=1e-12)
    assert np.array_equal(N, Ns)                 # hard labels: exact integers


@settings(max_examples=40, deadline=None, derandomize=True)
@given(shapes, st.integers(0, 2 ** 31 - 1))
def test_hard_label_counts_are_exact_and_complementary(shape, seed):
    x, y = _case(shape, seed)
    K = y.shape[1]
    _, N = O.pool_sums(x, O.weights_complement(y))
    assert np.array_equal(N, np.round(N))
    assert np.array_equal(N[:K] + N[K:], np.full(K, float(x.shape[0] * x.shape[2] * x.shape[3])))
    assert np.array_equal(N[:K], y.sum(axis=(0, 2, 3)))


@settings(max_examples=30, deadline=None, derandomize=True)
@given(shapes, st.integers(0, 2 ** 31 - 1))
def test_pool_backward_is_the_adjoint(shape, seed):
    """<g, d mu(x)[dx]> == <pool_backward(g), dx> (finite difference of the prototypes along a random direction)."""
    x, y = _case(shape, seed, hard=False)
    y = 0.1 + 0.8 * y                             # keep every weight sum away from 0
    w = O.weights_complement(y)
    rng = np.random.default_rng(seed + 7)
    g = rng.standard_normal((w.shape[1], x.shape[1]))
    dx = rng.standard_normal(x.shape)
    S, N = O.pool_sums(x, w)
    gx, _ = O.pool_backward(x, w, g, S, N)
    eps = 1e-6
    mu_p = O.prototypes_from_sums(*O.pool_sums(x + eps * dx, w))
    mu_m = O.prototypes_from_sums(*O.pool_sums(x - eps * dx, w))
    lhs = float((g * (mu_p - mu_m) / (2 * eps)).sum())
    rhs = float((gx * dx).sum())
    assert abs(lhs - rhs) <= 1e-6 * max(1.0, abs(rhs))


# ---- TransNorm (8(f) rank 4): the laws tests/test_gpu_transnorm.py relies on at full size
# at least 8 values per (half, channel): with 2-3 values the variance can be ~eps and the central differences below lose their footing
tn_shapes = st.tuples(st.integers(2, 6), st.integers(1, 6), st.integers(3, 5), st.integers(3, 5))


@settings(max_examples=40, deadline=None, derandomize=True)
@given(tn_shapes, st.integers(0, 2 ** 31 - 1))
def test_transnorm_normalises_each_half_and_alpha_sums_to_c(shape, seed):
    B, C, H, W = shape
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, C, H, W)) * rng.uniform(0.5, 3.0, (1, C, 1, 1)) + rng.standard_normal((1, C, 1, 1))
    gamma, beta = rng.uniform(0.5, 1.5, C), rng.standard_normal(C)
    eps = 1e-5
    fw = O.transnorm_train(x, gamma, beta, eps)
    assert np.isclose(fw["alpha"].sum(), C, rtol=1e-12)                 # alpha = C prob / sum(prob)
    h = B // 2
    z = fw["y"] / (1.0 + fw["alpha"])[None, :, None, None]
    for d, sl in enumerate((slice(0, h), slice(h, B))):
        m = z[sl].mean(axis=(0, 2, 3))
        v = z[sl].var(axis=(0, 2, 3))
        assert np.allclose(m, beta, atol=1e-9)                           # every half is centred on beta ...
        assert np.allclose(v, gamma ** 2 * fw["var_b"][d] / (fw["var_b"][d] + eps), rtol=1e-9, atol=1e-12)   # ... with scale gamma
    # permuting the samples inside each half or the pixels changes nothing but the order of the outputs
    pb = np.concatenate([rng.permutation(h), h + rng.permutation(B - h)])
    pp = rng.permutation(H * W)
    xp = x[pb].reshape(B, C, -1)[:, :, pp].reshape(x.shape)
    fp = O.transnorm_train(xp, gamma, beta, eps)
    assert np.allclose(fp["y"], fw["y"][pb].reshape(B, C, -1)[:, :, pp].reshape(x.shape), rtol=1e-10, atol=1e-10)
    assert np.allclose(fp["alpha"], fw["alpha"], rtol=1e-10)


@settings(max_examples=30, deadline=None, derandomize=True)
@given(tn_shapes, st.integers(0, 2 ** 31 - 1))
def test_transnorm_backward_is_the_adjoint(shape, seed):
    """<gy, J dx> == <J^T gy, dx> with alpha held constant, by central differences of the forward with frozen alpha."""
    B, C, H, W = shape
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, C, H, W)) * 1.5 + rng.standard_normal((1, C, 1, 1))
    gamma = rng.uniform(0.5, 1.5, C)
    gy, dx = rng.standard_normal(x.shape), rng.standard_normal(x.shape)
    gx, gw, gb = O.transnorm_train_backward(x, gamma, gy)
    alpha = O.transnorm_train(x, gamma, None)["alpha"]

    def frozen(xx, gg):
        fw = O.transnorm_train(xx, gg, None)
        return fw["y"] / (1.0 + fw["alpha"])[None, :, None, None] * (1.0 + alpha)[None, :, None, None]
    e = 1e-6
    jv = (frozen(x + e * dx, gamma) - frozen(x - e * dx, gamma)) / (2 * e)
    assert np.isclose((gy * jv).sum(), (gx * dx).sum(), rtol=2e-5, atol=1e-7)
    dg = rng.standard_normal(C)
    jg = (frozen(x, gamma + e * dg) - frozen(x, gamma - e * dg)) / (2 * e)
    assert np.isclose((gy * jg).sum(), (gw * dg).sum(), rtol=2e-5, atol=1e-7)
    assert np.allclose(gb, (gy * (1.0 + alpha)[None, :, None, None]).sum(axis=(0, 2, 3)), rtol=1e-10)
