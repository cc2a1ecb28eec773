"""B200-native batched legged-robot state estimator (orientation EKF + MHE hot path).

Drop-in for the estimator classes of well-robotics/Decentralized_EKF_MHE
(``DecentralizedEstimation``, ``MHEproblem``, ``orien_ekf``) on the data-parallel path only:
stepping many independent estimator instances per call on sm_100a.  See DESIGN.md.
"""
__all__ = ["synth", "params", "estimator", "sharding"]
