"""spatialaudiogen_b200: B200-native inference hot path of spatialaudiogen."""
from .definitions import *          # noqa: F401,F403
from .model import SptAudioGen, SptAudioGenParams  # noqa: F401
