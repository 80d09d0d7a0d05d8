"""Repo-root shim so that reference user scripts run unchanged:

    import arteryfe as af
    from arteryfe.artery import Artery
    from arteryfe.utils import *

Every name resolves to ``bloodflow_b200.arteryfe`` (the B200 implementation)."""
import sys

import bloodflow_b200.arteryfe as _impl
from bloodflow_b200.arteryfe import *          # noqa: F401,F403
from bloodflow_b200.arteryfe import ArteryNetwork, Artery, ParamParser, utils  # noqa: F401

for _name in ('artery', 'artery_network', 'param_parser', 'utils'):
    sys.modules[__name__ + '.' + _name] = getattr(_impl, _name)
