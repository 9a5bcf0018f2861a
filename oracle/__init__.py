"""CPU oracle for the dasp hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker (or as the timed CPU
baseline), never as a compute path of ``dasp_pytorch_b200``.
"""

from oracle.dasp_oracle import (  # noqa: F401
    gain,
    distortion,
    stereo_widener,
    stereo_panner,
    stereo_bus,
    biquad_section,
    eq_sections,
    sos_frequency_sampling,
    sos_recursion_truth,
    parametric_eq,
    compressor,
    expander,
    one_pole_recursion_truth,
    octave_filterbank,
    noise_shaped_reverberation,
    reverb_noise,
)
