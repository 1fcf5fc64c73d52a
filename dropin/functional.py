"""`from functional import modetqkrpb_cu` of ModeT-cu/models.py resolved to the sm_100a twins of modet_fw / modet_bw."""
from smilecode_b200.functional import ModeTFunction, modetqkrpb_cu  # noqa: F401
