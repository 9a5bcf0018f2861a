"""dasp_pytorch_b200: B200-native (sm_100a) kernels behind dasp_pytorch's functional audio processors.

Drop-in surface: the same names the reference package exports (``dasp_pytorch/__init__.py``) for the hot
path -- ``gain``, ``distortion``, ``parametric_eq``, ``compressor``, ``noise_shaped_reverberation`` and the
``Processor`` classes, ``stereo_bus``, ``stereo_panner``, ``stereo_widener`` -- plus ``expander`` (stubbed upstream).
"""
from dasp_pytorch_b200 import functional  # noqa: F401
from dasp_pytorch_b200.functional import (  # noqa: F401
    compressor,
    distortion,
    expander,
    gain,
    noise_shaped_reverberation,
    parametric_eq,
    stereo_bus,
    stereo_panner,
    stereo_widener,
)
from dasp_pytorch_b200.modules import (  # noqa: F401
    Compressor,
    Distortion,
    Expander,
    Gain,
    NoiseShapedReverb,
    ParametricEQ,
    Processor,
)
