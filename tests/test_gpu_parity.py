"""Parity of the CUDA path (through the C ABI) against the oracle, the reference fixtures and -- on the
same GPU -- the op-for-op eager restatement of the reference.  All tests need a B200 (``-m gpu``)."""
import numpy as np
import pytest
import torch

import uda_clr_b200 as clr
from oracle import clr_oracle as O
from oracle import clr_torch_port as TP
from uda_clr_b200 import synth
from _util import TOL_GRAD, TOL_PROTO, golden, relerr

pytestmark = pytest.mark.gpu
G = golden()
DEV = "cuda"


def cu(a, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(DEV)
    return t.requires_grad_(grad)


def stack(protos):
    return torch.cat([p.reshape(1, -1) for p in protos], 0).detach().cpu().numpy()


@pytest.mark.parametrize("case", ["hard_ragged", "hard_c305", "soft", "hard_big"])
def test_gen_prototype_vs_reference_fixture(case):
    c = G[case]
    soft = "grad_pred" in c
    pred, feat = cu(c["in_pred"], soft), cu(c["in_feat"], True)
    out = clr.gen_prototype(pred, feat)
    assert len(out) == 4 and all(o.shape == (1, feat.shape[1], 1, 1) and o.is_cuda for o in out)
    assert relerr(stack(out), c["out_protos"]) < TOL_PROTO
    seeds = cu(c["seed_g"])
    sum((o.reshape(-1) * s).sum() for o, s in zip(out, seeds)).backward()
    assert relerr(feat.grad.cpu().numpy(), c["grad_feat"]) < TOL_GRAD
    if soft:
        assert relerr(pred.grad.cpu().numpy(), c["grad_pred"]) < TOL_GRAD


@pytest.mark.parametrize("shape", [(8, 256, 128, 128, 2), (2, 305, 128, 128, 2), (3, 37, 24, 20, 2),
                                   (2, 16, 9, 7, 2), (2, 64, 32, 32, 4), (1, 40, 16, 16, 8), (2, 24, 16, 16, 3),
                                   (1, 12, 8, 8, 1)])
def test_gen_prototype_hard_vs_oracle(shape):
    B, C, H, W, K = shape
    g = torch.Generator().manual_seed(B * 1000 + C)
    y = synth.nested_ellipse_labels(B, K, H, W, g)
    x = synth.class_shifted_features(y, C, g)
    feat = x.to(DEV).requires_grad_(True)
    out = clr.gen_prototype(y.to(DEV), feat)
    S, N = O.pool_sums(x.numpy(), O.weights_complement(y.numpy()))
    assert relerr(stack(out), O.prototypes_from_sums(S, N)) < TOL_PROTO
    # integer work: the 2K pixel counts are bit-exact
    sums = clr.ops.pool_sums(feat.detach(), y.to(DEV), 0, K).cpu().numpy()
    assert np.array_equal(sums[:, C].astype(np.float64), N)
    seeds = torch.randn(2 * K, C, generator=g)
    sum((o.reshape(-1) * s).sum() for o, s in zip(out, seeds.to(DEV))).backward()
    gx, _ = O.gen_prototype_backward(y.numpy(), x.numpy(), seeds.numpy())
    assert relerr(feat.grad.cpu().numpy(), gx) < TOL_GRAD


def test_gen_prototype_soft_vs_eager_reference_on_gpu():
    """Same GPU, same inputs: our op vs the op-for-op eager restatement of utils/Utils.py:108-131."""
    b = synth.make_batch(B=4, C=96, H=64, W=64, K=2, image_res=False, seed=77)
    pred = torch.sigmoid(b.oT_before)
    p1, f1 = pred.to(DEV).requires_grad_(True), b.xt.to(DEV).requires_grad_(True)
    p2, f2 = pred.to(DEV).requires_grad_(True), b.xt.to(DEV).requires_grad_(True)
    ours, ref = clr.gen_prototype(p1, f1), TP.gen_prototype(p2, f2)
    assert relerr(stack(ours), stack(ref)) < TOL_PROTO
    seeds = torch.randn(4, 96, generator=torch.Generator().manual_seed(1)).to(DEV)
    sum((o.reshape(-1) * s).sum() for o, s in zip(ours, seeds)).backward()
    sum((o.reshape(-1) * s).sum() for o, s in zip(ref, seeds)).backward()
    assert relerr(f1.grad.cpu().numpy(), f2.grad.cpu().numpy()) < TOL_GRAD
    assert relerr(p1.grad.cpu().numpy(), p2.grad.cpu().numpy()) < TOL_GRAD


def test_src_trg_fixture():
    c = G["src_trg"]
    out = clr.gen_prototype_src_trg(cu(c["in_pred_s"]), cu(c["in_feat_s"]), cu(c["in_pred_t"]), cu(c["in_feat_t"]))
    assert relerr(stack(out), c["out_protos"]) < TOL_PROTO


def test_empty_class_gives_nan_like_reference():
    y = torch.zeros(1, 2, 8, 8)
    y[:, 1, 2:6, 2:6] = 1.0
    x = torch.randn(1, 8, 8, 8)
    out = stack(clr.gen_prototype(y.to(DEV), x.to(DEV)))
    assert np.isnan(out[0]).all() and not np.isnan(out[1:]).any()


def test_noncontiguous_and_misaligned_inputs():
    """Strided views are made contiguous; a base pointer that is only 4-byte aligned takes the scalar path."""
    g = torch.Generator().manual_seed(5)
    y = synth.nested_ellipse_labels(2, 2, 16, 16, g)
    x = synth.class_shifted_features(y, 10, g)
    ref = O.gen_prototype(y.numpy(), x.numpy())
    xp = torch.zeros(2, 10, 16, 20).to(DEV)
    xp[..., 2:18] = x.to(DEV)
    assert relerr(stack(clr.gen_prototype(y.to(DEV), xp[..., 2:18])), ref) < TOL_PROTO
    flat = torch.zeros(x.numel() + 1, device=DEV)
    flat[1:] = x.to(DEV).reshape(-1)
    xm = flat[1:].view(2, 10, 16, 16)
    assert xm.data_ptr() % 16 == 4
    assert relerr(stack(clr.gen_prototype(y.to(DEV), xm)), ref) < TOL_PROTO


def test_linearity_and_shard_sum_property_full_size():
    """Size-independent properties at the bench size (B=8, C=256, 128x128): pooling is linear in the
    features, and shard-and-sum of the packed sums equals the whole batch (the all-reduce decomposition)."""
    b = synth.make_batch(B=8, C=256, H=128, W=128, K=2, image_res=False, seed=1234)
    y, x = b.ys.to(DEV), b.xs.to(DEV)
    whole = clr.ops.pool_sums(x, y, 0, 2).double()
    parts = sum(clr.ops.pool_sums(x[i:i + 2].contiguous(), y[i:i + 2].contiguous(), 0, 2).double() for i in range(0, 8, 2))
    assert relerr(parts.cpu().numpy(), whole.cpu().numpy()) < 1e-6
    assert torch.equal(parts[:, 256], whole[:, 256])          # counts: exact
    assert float(whole[0, 256] + whole[2, 256]) == 8 * 128 * 128  # obj + bck = all pixels
    a = clr.ops.pool_sums(2.5 * x, y, 0, 2).double()
    assert relerr(a[:, :256].cpu().numpy(), 2.5 * whole[:, :256].cpu().numpy()) < 1e-6
    # run-to-run bit stability (fixed-order reduction, no atomics)
    again = clr.ops.pool_sums(x, y, 0, 2).double()
    assert torch.equal(again, whole)


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("shape", [(8, 256, 128, 128, 2), (2, 305, 64, 64, 2), (2, 64, 32, 32, 4), (1, 40, 16, 16, 8)])
def test_both_pooling_kernels_agree_with_oracle(impl, shape):
    """The 128-bit LDG kernel (pool_impl=1, the default) and the bulk-TMA ring kernel (pool_impl=2) are the same
    decomposition with different data paths: both must meet the prototype tolerance and give bit-exact counts."""
    from uda_clr_b200 import _lib
    lib = _lib.load()
    B, C, H, W, K = shape
    g = torch.Generator().manual_seed(77 + C)
    y = synth.nested_ellipse_labels(B, K, H, W, g)
    x = synth.class_shifted_features(y, C, g)
    S, N = O.pool_sums(x.numpy(), O.weights_complement(y.numpy()))
    try:
        _lib.check(lib.clr_set_tunable(b"pool_impl", impl), "pool_impl")
        sums = clr.ops.pool_sums(x.to(DEV), y.to(DEV), 0, K).cpu().numpy()
    finally:
        lib.clr_set_tunable(b"pool_impl", 0)
    assert np.array_equal(sums[:, C].astype(np.float64), N)
    assert relerr(sums[:, :C], S) < TOL_PROTO


def test_two_to_the_31_elements_feature_map():
    """The sweep corner B=64, C=512, 256x256 is exactly 2^31 floats (8.6 GB): int32 element offsets overflow there.
    Size-independent checks: exact counts, shard-and-sum == whole, and the adjoint at the far end of the address space."""
    free, _ = torch.cuda.mem_get_info()
    if free < 30 * 2 ** 30:
        pytest.skip("needs ~26 GB of free device memory")
    B, C, H, W, K = 64, 512, 256, 256, 2
    g = torch.Generator().manual_seed(5)
    y = synth.nested_ellipse_labels(B, K, H, W, g).to(DEV)
    x = torch.randn(B, C, H, W, device=DEV)
    assert x.numel() == 2 ** 31
    whole = clr.ops.pool_sums(x, y, 0, K).double()
    halves = sum(clr.ops.pool_sums(x[i:i + 32], y[i:i + 32].contiguous(), 0, K).double() for i in (0, 32))
    assert torch.equal(halves[:, C], whole[:, C])                        # counts exact
    assert float(whole[0, C] + whole[K, C]) == B * H * W                 # obj + bck = all pixels (< 2^24: exact in fp32)
    assert float(whole[0, C]) == float(y[:, 0].sum())
    assert relerr(halves.cpu().numpy(), whole.cpu().numpy()) < 1e-6
    # last sample, last channels: direct fp64 evaluation of S_r[c] for a probe
    probe = torch.stack([(x[:, C - 1].double() * y[:, k].double()).sum() for k in range(K)])
    assert relerr(whole[:K, C - 1].cpu().numpy(), probe.cpu().numpy()) < 1e-5
    # adjoint: grad[b, c, p] = sum_r g[r][c] / N_r * w_r[b, p]; check the very last plane
    gmat = torch.randn(2 * K, C, device=DEV)
    grad = clr.ops.pool_backward_feat(y, 0, K, (B, C, H, W), gmat, whole.float(), 1.0)
    N = whole[:, C]
    wlast = torch.cat([y[B - 1], 1.0 - y[B - 1]], 0).double()            # [2K, H, W] complement rows
    ref = torch.einsum("r,rhw->hw", (gmat[:, C - 1].double() / N), wlast)
    assert relerr(grad[B - 1, C - 1].cpu().numpy(), ref.cpu().numpy()) < TOL_GRAD
    del x, grad
    torch.cuda.empty_cache()


def test_plain_c_host_runs_against_the_library(tmp_path):
    """examples/clr_host.c: a C99 program that links libclr_b200.so + cudart only, runs pooling / prototypes / the
    adjoint through the C ABI and checks them against a closed form in double on the host."""
    import os
    import shutil
    import subprocess
    from uda_clr_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("gcc") or not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("gcc / CUDA toolkit not available")
    exe = str(tmp_path / "clr_host")
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                        os.path.join(root, "examples", "clr_host.c"), "-o", exe, "-L", libdir, "-lclr_b200",
                        "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "clr_host ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("K", [1, 2, 4, 5, 8])
def test_pool_cta_sizes_agree_with_oracle(K):
    """The 128-bit pooling kernel is built for 256- and 128-thread CTAs (same chunk, same partial layout; 128 is the
    default for more than 8 weight rows): both must reproduce the oracle's sums, and exact integer counts."""
    from uda_clr_b200 import _lib, ops
    lib = _lib.load()
    g = torch.Generator().manual_seed(40 + K)
    B, C, H, W = 3, 37, 48, 64
    y = synth.nested_ellipse_labels(B, K, H, W, g)
    x = synth.class_shifted_features(y, C, g)
    S, N = O.pool_sums(x.numpy(), O.weights_complement(y.numpy()))
    want = np.concatenate([S, N[:, None]], axis=1)
    got = {}
    try:
        for nt in (256, 128):
            _lib.check(lib.clr_set_tunable(b"pool_threads", nt), "pool_threads")
            got[nt] = ops.pool_sums(x.to(DEV), y.to(DEV), _lib.CLR_W_COMPLEMENT, K).cpu().numpy()
    finally:
        lib.clr_set_tunable(b"pool_threads", 0)
    for nt, s in got.items():
        assert relerr(s[:, :-1], want[:, :-1]) < TOL_PROTO
        assert np.array_equal(s[:, -1], want[:, -1])          # hard labels: pixel counts are exact integers
    assert relerr(got[128], got[256]) < 2e-6
