"""uda_clr_b200 -- B200-native category-level-regularisation (CLR) hot path of fengweie/UDA_CLR.

Public surface = the reference's own function names (``utils/Utils.py``), backed by hand-written
sm_100a CUDA kernels behind a C ABI (``include/clr_b200.h``, ``libclr_b200.so``):

    from uda_clr_b200 import gen_prototype, gen_prototype_retrify, ...
    uda_clr_b200.patch_reference()      # rebind the names inside an imported reference tree

Importing the package never touches CUDA; the first op call loads the library and raises if it is
missing (no CPU fallback).
"""
from .ops import (adaptation_factor, bmm_prototypes, dice_from_counts, distance_weight, feat_prototype_distance, gen_prototype,  # noqa: F401
                  gen_prototype_retrify, gen_prototype_src_trg, gen_prototype_src_trg_retrify,
                  get_prototype_weight, mc_statistics, pixel_acc_from_counts, retrify_weights, seg_loss,
                  uncertainty_map, validation_counts,
                  update_objective_single_vector, weighted_prototypes, nearest_labels, MCAccumulator)
from .step import CLRPlan, CLRStep, CLRStepError, CLRStepOutput, consistency_threshold, sigmoid_rampup  # noqa: F401
from .offline import OfflinePrototypes, offline_masks  # noqa: F401
from . import dist, ops  # noqa: F401
from .patch import patch_reference, patch_transnorm, unpatch_reference  # noqa: F401
from .transnorm import TransNorm1d, TransNorm2d, trans_norm  # noqa: F401

__version__ = "0.1.0"
