"""Put this directory first on sys.path (or copy this file over ModeT/models.py) and the reference's
`from models import ModeT` (ModeT/train.py:14, ModeT/infer.py:12) and
`from models import ModeT_cu` (ModeT-cu/train.py:14) resolve to the B200-native implementation."""
from smilecode_b200.models import *  # noqa: F401,F403
from smilecode_b200.models import ModeT, ModeT_cu  # noqa: F401
