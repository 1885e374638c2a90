"""fsb200 -- Python side of libfsb200.so, the B200-native hot path of freesound-classification."""
from ._lib import EXPORTS, LIB_PATH, build, check, lib  # noqa: F401
