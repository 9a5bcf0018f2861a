"""dasp_pytorch_b200: B200-native kernels behind dasp_pytorch's functional audio processors."""
from dasp_pytorch_b200.functional import (  # noqa: F401
    gain,
    distortion,
    compressor,
    expander,
    parametric_eq,
    noise_shaped_reverberation,
)
from dasp_pytorch_b200 import functional  # noqa: F401
