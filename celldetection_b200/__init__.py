"""B200-native Contour Proposal Network inference -- drop-in for the CPN hot path of FZJ-INM1-BDA/celldetection.

``import celldetection_b200 as cd`` exposes the names the path uses: ``cd.models.CPN / CpnU22 / CpnResNet18FPN /
CpnResNeXt101UNet``, ``cd.ops.cpn.*``, ``cd.fetch_model`` / ``cd.load_model`` and ``cd.cpn_inference`` /
``cd.apply_model``.  Everything executes in hand-written sm_100a CUDA behind the C ABI of ``include/cpn_b200.h``.
"""
from . import _lib
from . import ops
from . import models
from . import data
from .utils import load_model, fetch_model, save_fetchable_model, synth_state_dict, calibrate_heads_, to_h5, from_h5
from .inference import get_tiling_slices, apply_model, cpn_inference
from .preprocessing import preprocess

__version__ = '0.1.0'
