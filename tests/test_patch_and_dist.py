"""Host-side logic that needs no GPU: the reference patch seam and the data-parallel decomposition
(world_size-2 gloo run: shard -> packed sums -> all-reduce -> finalize == whole batch)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import uda_clr_b200 as clr
from oracle import clr_oracle as O
from oracle import ref_import
from uda_clr_b200 import synth


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_patch_reference_rebinds_star_import_copies():
    import datetime
    import types
    U = ref_import.load_utils()
    pytz = types.ModuleType("pytz")
    pytz.timezone = lambda n: datetime.timezone.utc
    saved = sys.modules.get("pytz")
    sys.modules["pytz"] = pytz
    try:
        import train_process.Trainer_prototype_full as TPF   # star-imports utils.Utils (:16)
    finally:
        if saved is not None:
            sys.modules["pytz"] = saved
    orig = U.gen_prototype
    assert TPF.gen_prototype is orig
    report = clr.patch_reference()
    try:
        assert U.gen_prototype is clr.gen_prototype
        assert TPF.gen_prototype is clr.gen_prototype
        assert TPF.gen_prototype_retrify is clr.gen_prototype_retrify
        assert "train_process.Trainer_prototype_full" in report["gen_prototype_retrify"]
        assert U.adaptation_factor(3.0) == clr.adaptation_factor(3.0)
        with pytest.raises(RuntimeError):
            clr.patch_reference()
    finally:
        clr.unpatch_reference()
    assert U.gen_prototype is orig and TPF.gen_prototype is orig


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_patch_transnorm_rebuilds_deeplab_with_our_module():
    """``DeepLab(sync_bn=False)`` (--use_TN, networks/deeplabv3.py:17-27) built after ``patch_transnorm()`` carries
    ``TransNorm2d`` everywhere the reference puts its TransNorm, with an identical state-dict layout."""
    ref_import.ref_transnorm_class()
    import networks.deeplabv3 as dl
    from networks.backbone import mobilenet
    from uda_clr_b200.transnorm import TransNorm2d
    saved = mobilenet.MobileNetV2._load_pretrained_model
    mobilenet.MobileNetV2._load_pretrained_model = lambda self: None        # no network: random init
    try:
        ref_model = dl.DeepLab(num_classes=2, backbone="mobilenet", output_stride=16, sync_bn=False, freeze_bn=False)
        patched = clr.patch_transnorm()
        try:
            assert "networks.deeplabv3" in patched
            ours = dl.DeepLab(num_classes=2, backbone="mobilenet", output_stride=16, sync_bn=False, freeze_bn=False)
        finally:
            clr.unpatch_reference()
        assert dl.BatchNorm2d is not TransNorm2d
    finally:
        mobilenet.MobileNetV2._load_pretrained_model = saved
    n_ref = sum(1 for m in ref_model.modules() if type(m).__name__ == "BatchNorm2d" and hasattr(m, "running_mean_source"))
    n_ours = sum(1 for m in ours.modules() if isinstance(m, TransNorm2d))
    assert n_ref == n_ours and n_ours > 30
    sd_ref, sd_ours = ref_model.state_dict(), ours.state_dict()
    assert list(sd_ref) == list(sd_ours)
    assert all(sd_ref[k].shape == sd_ours[k].shape and sd_ref[k].dtype == sd_ours[k].dtype for k in sd_ref)
    ours.load_state_dict(sd_ref)        # checkpoints of the reference load into the patched model


def test_drop_in_signatures_match_reference_source():
    """Positional parameter names of the drop-ins equal the reference's (utils/Utils.py:86-311)."""
    import inspect
    expect = {
        "gen_prototype": ["pred_oS", "xs_feature"],
        "gen_prototype_src_trg": ["pred_oS", "xs_feature", "pred_oT", "xt_feature"],
        "gen_prototype_retrify": ["oT_before", "xt_feature", "preds", "features", "T", "stride"],
        "gen_prototype_src_trg_retrify": ["pred_oS", "xs_feature", "oT_before", "xt_feature", "preds", "features",
                                          "T", "stride"],
        "get_prototype_weight": ["feat", "class_num", "prototype"],
        "adaptation_factor": ["m"],
    }
    for name, params in expect.items():
        assert list(inspect.signature(getattr(clr, name)).parameters) == params
    if ref_import.available():
        U = ref_import.load_utils()
        for name, params in expect.items():
            assert list(inspect.signature(getattr(U, name)).parameters) == params


def test_shard_bounds_cover_batch():
    for n in (1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [clr.dist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        clr.dist.shard_bounds(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clr.dist.enable()
        assert clr.dist.enabled() and clr.dist.world_size() == world and clr.dist.grad_scale() == float(world)
        b = synth.make_batch(B=4, C=9, H=8, W=8, K=2, image_res=False, seed=77)
        ys, xs, xt = clr.dist.shard_batch([b.ys, b.xs, b.xt], rank, world)
        wt = torch.sigmoid(clr.dist.shard_batch([b.oT_before], rank, world)[0])
        # per-rank packed sums (what clr_pool_fwd produces on the GPU): [2 domains][2K][C+1]
        Ss, Ns = O.pool_sums(xs.numpy(), O.weights_complement(ys.numpy()))
        St, Nt = O.pool_sums(xt.numpy(), O.weights_complement(wt.numpy()))
        packed = torch.tensor(np.stack([np.concatenate([Ss, Ns[:, None]], 1), np.concatenate([St, Nt[:, None]], 1)]),
                              dtype=torch.float64)
        clr.dist.all_reduce_sums(packed)            # THE exchange of the path
        g = packed.numpy()
        glob = dict(Ss=g[0, :, :-1], Ns=g[0, :, -1], St=g[1, :, :-1], Nt=g[1, :, -1])
        o = O.clr_step(xs.numpy(), ys.numpy(), xt.numpy(), wt.numpy(), w_intra=0.1, global_sums=glob,
                       grad_scale=clr.dist.grad_scale())
        out_q.put((rank, o["intra"], o["Ps"], o["gxs"], o["gxt"]))
    finally:
        clr.dist.disable()
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_allreduce_equals_whole_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = synth.make_batch(B=4, C=9, H=8, W=8, K=2, image_res=False, seed=77)
    wt = torch.sigmoid(b.oT_before)
    whole = O.clr_step(b.xs.numpy(), b.ys.numpy(), b.xt.numpy(), wt.numpy(), w_intra=0.1)
    for rank, intra, Ps, gxs, gxt in res:
        assert abs(intra - whole["intra"]) < 1e-12
        assert np.allclose(Ps, whole["Ps"], rtol=1e-12, atol=0)
    # DDP averages gradients: (1/G) * sum_ranks (G * local) must equal the single-process gradient
    gxs = np.concatenate([r[3] for r in res], 0) / world
    gxt = np.concatenate([r[4] for r in res], 0) / world
    assert np.allclose(gxs, whole["gxs"], rtol=1e-10, atol=1e-18)
    assert np.allclose(gxt, whole["gxt"], rtol=1e-10, atol=1e-18)


@pytest.mark.reference
def test_real_trainer_train_epoch_stock_vs_port_at_the_patch_seam():
    """Whole-trainer integration: the UNMODIFIED ``Trainer_prototype_full.Trainer.train_epoch`` runs two
    steps (MobileNetV2 DeepLab, 512x512, batch 2) stock and with ``gen_prototype`` / ``gen_prototype_retrify`` rebound on
    the trainer module -- the seam ``patch_reference()`` uses -- to the eager port; every prototype either run's ops
    returned must be identical.  ~100 s of CPU: opt-in with CLR_RUN_TRAINER_TEST=1 (result of the last run:
    profiles/r02_trainer_harness_cpu.json).  The GPU form (patched = the CUDA ops) is tests/test_gpu_integration.py."""
    if os.environ.get("CLR_RUN_TRAINER_TEST") != "1":
        pytest.skip("opt-in (CLR_RUN_TRAINER_TEST=1): ~100 s")
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
    import trainer_harness as TH
    stock = TH.run_train_epoch(ref_import.REFERENCE_ROOT, False, steps=2, batch_size=2)
    port = TH.run_train_epoch(ref_import.REFERENCE_ROOT, "port", steps=2, batch_size=2)
    assert len(stock["calls"]) == len(port["calls"]) == 4
    for a, b in zip(stock["calls"], port["calls"]):
        assert a["name"] == b["name"]
        for x, y in zip(a["protos"], b["protos"]):
            assert np.array_equal(x, y)
    assert stock["running_intra"] == port["running_intra"]


def test_step_state_dict_is_a_snapshot_and_load_validates():
    """``state_dict()`` must not hand out the live EMA tensors (later steps update them in place) and
    ``load_state_dict`` must reject state that does not fit the step before raw pointers reach the kernels."""
    step = clr.CLRStep(K=2)
    assert step.state_dict()["stored_s"] is None and step.state_dict()["first_s"] is True
    step.stored_s, step.stored_t = torch.ones(4, 7), torch.full((4, 7), 2.0)
    step.first_s = step.first_t = False
    sd = step.state_dict()
    step.stored_s.add_(1.0)                                   # what a later step does, in place
    assert float(sd["stored_s"][0, 0]) == 1.0                 # the snapshot did not move
    other = clr.CLRStep(K=2)
    other.load_state_dict(sd)
    assert other.first_s is False and torch.equal(other.stored_s, sd["stored_s"]) and other.stored_s is not sd["stored_s"]
    for bad in (dict(sd, stored_s=torch.ones(3, 7)),                      # wrong 2K
                dict(sd, stored_t=torch.ones(4, 8)),                      # shapes disagree
                dict(sd, stored_s=torch.ones(4, 7, dtype=torch.float64)),
                dict(sd, stored_t=None)):
        with pytest.raises(ValueError):
            clr.CLRStep(K=2).load_state_dict(bad)


def test_dist_local_context_restores_the_exchange_settings():
    saved = dict(clr.dist._STATE)
    try:
        clr.dist._STATE.update(enabled=True, peer=True, grad_scale=4.0)
        with clr.dist.local():
            assert not clr.dist.enabled() and not clr.dist.peer_enabled() and clr.dist.world_size() == 1
        assert clr.dist._STATE["enabled"] and clr.dist._STATE["peer"] and clr.dist._STATE["grad_scale"] == 4.0
    finally:
        clr.dist._STATE.update(saved)
