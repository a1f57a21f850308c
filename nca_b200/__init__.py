"""Import alias: the product package lives in ``video-stylization-with-nca_b200/`` (not a valid Python
identifier), so ``nca_b200`` extends its ``__path__`` there and re-exports the public names."""
import os as _os

_PKG = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "video-stylization-with-nca_b200")
__path__.append(_PKG)

from ._lib import NcaError, lib_path, load_library  # noqa: E402,F401
from . import functional, parallel, trainer, video  # noqa: E402,F401
from .trainer import NormalizedAdam, TensorSamplePool, pool_gather, pool_scatter, overflow_loss  # noqa: E402,F401
from .video import FrameStylizer  # noqa: E402,F401
from .dynca_ec import DyNCA as DyNCA_EC, CPE2D  # noqa: E402,F401
from .dynca_cd import DyNCA as DyNCA_CD, EdgeExtractor  # noqa: E402,F401
from .nca import ConditionedNCA, UpdateNet, ImageEncoder  # noqa: E402,F401
