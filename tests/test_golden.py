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
