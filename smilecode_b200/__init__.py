"""smilecode_b200: the ModeT deformable-registration hot path (ZAX130/SmileCode) as hand-written
sm_100a CUDA kernels behind a C ABI, with drop-in `models.py` classes on top.

    from smilecode_b200.models import ModeT          # same signature / state_dict as the reference
    python -m smilecode_b200.build                   # builds libsmilecode_b200.so in-tree (nvcc)
"""
__version__ = "0.1.0"
