"""Drop-in mirror of the reference's ``arteryfe`` package namespace
(arteryfe/__init__.py:1-3) on top of the B200 CUDA library."""
from .utils import *                       # noqa: F401,F403
from .param_parser import ParamParser      # noqa: F401
from .artery_network import ArteryNetwork  # noqa: F401
from .artery import Artery                 # noqa: F401
