"""stair_b200 — B200 (sm_100a) implementation of STAIR's video_nmn ModuleNet hot path.

Public surface (mirrors video_nmn/module_net.py + video_nmn/modules.py of the reference):

    from stair_b200 import VideoNMN, collate, NAME_TO_MODULE
"""
from .layout import NARY as nary_mappings, MODULE_NAMES, WORDS_TO_KEEP, collate, collate_chunks, compile_layout, NMNBatch  # noqa: F401
from .nmn import VideoNMN  # noqa: F401
from .params import L2Normalize  # noqa: F401

from .modules import NAME_TO_MODULE  # noqa: F401,E402   name -> operator class, the reference's registration order (modules.py:446-465)

assert list(NAME_TO_MODULE) == list(MODULE_NAMES)
__version__ = '0.1.0'
