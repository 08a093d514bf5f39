"""libcluster_b200 -- B200-native variational E-step engine behind libcluster's
learnBGMM / learnVDP / learnGMC / learnDGMM entry points and the distributions.h
operator surface.  Native code: libcluster_b200/csrc (CUDA for sm_100a + C ABI);
this package is the thin ctypes mirror of the reference's Python module.
"""
from .api import (BGMM, DGMC, DGMM, F32, F64, GMC, PRIORVAL, SGMC, SPLITITER, VDP, CudaError, Dirichlet,  # noqa: F401
                  DomainError, Engine, FreeEnergyError, GaussWish, GDirichlet, InvalidArgument, NormGamma,
                  StickBreak, default_engine, host_mstep, learnBGMM, learnDGMC, learnDGMM, learnGMC, learnSGMC,
                  learnVDP, nccl_unique_id, packed_len, shard_rows)

__version__ = "0.1.0"
