"""rlitg_b200 -- B200 (sm_100a) kernels for RL-ITG's kinematics / cost hot path, behind the reference's call
signatures.  Everything computes in ``librlitg.so`` (hand-written CUDA, C ABI in ``include/rlitg.h``); there is
no CPU or PyTorch fallback -- importing a compute module without the built library raises."""
__version__ = "0.1.0"
