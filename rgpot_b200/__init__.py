"""rgpot_b200 -- B200 (sm_100a) CUDA engine for rgpot's LJ / LJCluster / CuH2 hot path.

Public surface (mirrors the reference's Python module, python/bind/bind_module.cpp:142-174, plus
the potentials the reference only exposes from C++):

    evaluate_lj(positions, atom_types, box) -> (energy, forces, variance)
    LJPot()(positions, atom_types, box)
    LJClusterPot, CuH2Pot, evaluate_cuh2, <pot>.forceBatch([...])

Everything is computed by librgpot_b200.so (include/rgpot_b200.h).  No CPU fallback.
"""
from .potentials import (CuH2Pot, EAMAlPot, FeHePot, LJClusterConfig, LJClusterPot, LJConfig, LJPot, MorseConfig, MorsePot, PotCaps, ZBLConfig, ZBLPot,
                         PotentialError, PotType, Reentrancy, evaluate_batch, evaluate_cuh2, evaluate_lj,
                         measure_copy_gbs, measure_fp64_tflops, measure_host_copy_gbs, neighbor_list, pinned_empty)

__version__ = "0.1.0"
__all__ = ["CuH2Pot", "EAMAlPot", "FeHePot", "LJClusterConfig", "LJClusterPot", "LJConfig", "LJPot", "MorseConfig", "MorsePot", "PotCaps", "ZBLConfig", "ZBLPot",
           "PotentialError", "PotType", "Reentrancy", "evaluate_batch", "evaluate_cuh2", "evaluate_lj",
           "measure_copy_gbs", "measure_fp64_tflops", "measure_host_copy_gbs", "neighbor_list", "pinned_empty", "__version__"]
