"""multitreegp_b200 -- B200-native (sm_100a) population fitness evaluation for MultiTreeGP.

Only the hot path of the reference is implemented here: the bytecode
encoder, the fused tree-interpreter / ODE-rollout kernels, the pointwise interpreter kernel and
the host-side mirror of the reference's evaluator / environment interface that feeds them.
"""
__version__ = "0.1.0"
