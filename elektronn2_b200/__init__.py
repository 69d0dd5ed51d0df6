"""elektronn2_b200 -- B200 (sm_100a) implementation of ELEKTRONN2's volumetric-CNN hot
path: Conv / UpConv / Pool / MFP forward+backward and tiled ``predict_dense``,
behind the reference's neuromancer node API.

    from elektronn2_b200 import neuromancer as nm      # same surface as elektronn2.neuromancer
    elektronn2_b200.install_as_elektronn2()            # lets unedited model files import `elektronn2`
"""
import sys
import types

from .config import config  # noqa: F401

__version__ = '0.1.0'


def install_as_elektronn2():
    """Register this package under the name ``elektronn2`` so the reference's example /
    config files (``from elektronn2 import neuromancer as nm``) run unedited."""
    from . import neuromancer
    shim = types.ModuleType('elektronn2')
    shim.neuromancer = neuromancer
    shim.config = config
    shim.__path__ = []
    sys.modules['elektronn2'] = shim
    sys.modules['elektronn2.neuromancer'] = neuromancer
    return shim
