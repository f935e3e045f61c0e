"""Oracle vs the committed reference fixtures (tests/golden/ref_golden.npz, recorded from the unmodified
reference by tests/golden/make_golden.py).  CPU only; runs everywhere."""
import numpy as np
import pytest

from oracle import clr_oracle as O
from _util import TOL_GRAD, TOL_PROTO, golden, relerr

G = golden()


@pytest.mark.parametrize("case", ["hard_ragged", "hard_c305", "soft", "hard_big"])
def test_gen_prototype_golden(case):
    c = G[case]
    mu = O.gen_prototype(c["in_pred"], c["in_feat"])
    assert relerr(mu, c["out_protos"]) < TOL_PROTO
    gx, gp = O.gen_prototype_backward(c["in_pred"], c["in_feat"], c["seed_g"])
    assert relerr(gx, c["grad_feat"]) < TOL_GRAD
    if "grad_pred" in c:
        assert relerr(gp, c["grad_pred"]) < TOL_GRAD
    else:
        # hard labels: the four pixel counts are exact integers
        _, N = O.pool_sums(c["in_feat"], O.weights_complement(c["in_pred"]))
        assert np.array_equal(N, np.round(N))


def test_src_trg_golden():
    c = G["src_trg"]
    mu = O.gen_prototype_src_trg(c["in_pred_s"], c["in_feat_s"], c["in_pred_t"], c["in_feat_t"])
    assert relerr(mu, c["out_protos"]) < TOL_PROTO


def test_retrify_golden():
    c = G["retrify"]
    T = int(c["in_T"])
    preds = c["in_preds_f16"].astype(np.float32)
    B = c["in_xt"].shape[0]
    o = O.gen_prototype_retrify(c["in_oT_before"], c["in_xt"], preds, T, B)
    assert relerr(o["std_map"], c["out_std_map"]) < 5e-6
    assert np.array_equal(o["mask_0"].astype(np.uint8), c["out_mask_0"])   # integer work: bit-exact
    assert np.array_equal(o["mask_1"].astype(np.uint8), c["out_mask_1"])
    assert relerr(o["protos"], c["out_protos"]) < TOL_PROTO
    gx, _ = O.pool_backward(c["in_xt"], o["w"], c["seed_g"])
    assert relerr(gx, c["grad_xt"]) < TOL_GRAD


def test_cosine_and_schedule_golden():
    c = G["cosine"]
    assert relerr(O.cosine_weight(c["in_feat"], c["in_proto"]), c["out_weight"]) < 1e-6
    ours = np.array([O.adaptation_factor(float(m)) for m in c["in_m"]])
    assert np.array_equal(ours, c["out_adaptation_factor"])


# ---- 8(f) rank 4: TransNorm fixtures recorded from the reference module (tests/golden/make_transnorm_golden.py)
from _util import TOL_TN, transnorm_golden  # noqa: E402

TN = transnorm_golden()


@pytest.mark.parametrize("case", sorted(TN))
def test_transnorm_golden(case):
    c = TN[case]
    w, b, gy = c["in_weight"], c["in_bias"], c["seed_gy"]
    run = {k: (np.zeros_like(w) if "mean" in k else np.ones_like(w)).astype(np.float64)
           for k in ("running_mean_source", "running_var_source", "running_mean_target", "running_var_target")}
    for step, key in enumerate(("in_x", "in_x2")):
        x = c[key]
        fw = O.transnorm_train(x, w, b)
        assert relerr(fw["y"], c["out_y%d" % step]) < TOL_TN
        gx, gw, gb = O.transnorm_train_backward(x, w, gy)
        assert relerr(gx, c["grad_x%d" % step]) < TOL_TN * 10     # the reference's own fp32 backward carries ~1e-5
        assert relerr(gw, c["grad_weight%d" % step]) < TOL_TN
        assert relerr(gb, c["grad_bias%d" % step]) < TOL_TN
        for d, dom in enumerate(("source", "target")):
            run["running_mean_" + dom] = O.transnorm_running(run["running_mean_" + dom], fw["mean"][d], 0.1)
            run["running_var_" + dom] = O.transnorm_running(run["running_var_" + dom], fw["var_u"][d], 0.1)
        for k, v in run.items():
            assert relerr(v, c["out_%s%d" % (k, step)]) < TOL_TN
    y = O.transnorm_eval(c["in_x"], w, b, run["running_mean_source"], run["running_var_source"],
                         run["running_mean_target"], run["running_var_target"])
    assert relerr(y, c["out_eval"]) < TOL_TN
