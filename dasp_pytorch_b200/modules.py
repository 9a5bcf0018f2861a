"""Normalised-parameter processors over the B200 kernels.

Same contract as the reference's ``dasp_pytorch/modules.py`` (@ c9ae0126): a ``Processor`` owns an ordered
``param_ranges`` dict, ``process_normalized(x, p)`` takes ``p`` in ``[0, 1]`` with shape
``(batch, num_params)`` (columns in ``param_ranges`` order), maps each column affinely onto its range and
calls ``process_fn(x, sample_rate, **params)`` with the parameter NAMES as keywords (reference
``modules.py:25-51``) -- which is why the keyword names of ``dasp_pytorch_b200.functional`` are part of the
drop-in contract.  ``process(x, *args)`` forwards positionally (``modules.py:53-54``).

Differences, all on the host side:
  * the range check of ``denormalize_param_dict`` (reference ``modules.py:83-84``) costs the reference two
    device->host synchronisations per parameter; here it is ONE fused check per call;
  * ``Distortion`` takes a ``sample_rate`` and uses the functional's real keyword ``drive_db``: the reference
    class cannot be called through ``process_normalized`` at all (it has no ``sample_rate`` attribute and
    names its parameter ``gain_db``).
"""
from __future__ import annotations

from typing import Dict

import torch

from dasp_pytorch_b200 import functional as _F
from dasp_pytorch_b200.functional import (
    compressor,
    distortion,
    expander,
    gain,
    noise_shaped_reverberation,
    parametric_eq,
)


def denormalize(norm_val, max_val, min_val):
    return (norm_val * (max_val - min_val)) + min_val


def normalize(val, min_val, max_val):
    return (val - min_val) / (max_val - min_val)


class Processor:
    """Base class: subclasses set ``sample_rate``, ``process_fn`` and ``param_ranges``."""

    sample_rate = None
    process_fn = None
    param_ranges: Dict[str, tuple] = {}

    @property
    def num_params(self) -> int:
        return len(self.param_ranges)

    # subclasses with a packed kernel entry point set this to (callable(x, sr, packed), the process_fn it mirrors)
    _packed_path = None

    def process_normalized(self, x: torch.Tensor, param_tensor: torch.Tensor):
        """Run the processor with parameters normalised to (0, 1), shape ``(batch, num_params)``.

        Fast path: when ``process_fn`` is still this package's kernel entry, the whole
        parameter handling is ONE range-check reduction and ONE affine kernel on the packed ``(batch, P)``
        tensor, which then goes to the kernels as is -- instead of the reference's per-parameter slicing,
        2 host syncs and ~3 tiny kernels per parameter (modules.py:56-91).  Same errors, same results.
        """
        if self._packed_path is not None and self.process_fn is self._packed_path[1] and param_tensor.is_cuda:
            if param_tensor.dim() != 2 or param_tensor.shape[1] != len(self.param_ranges):
                raise ValueError(
                    f"Parameter tensor has {param_tensor.shape[1] if param_tensor.dim() == 2 else '?'} parameters, "
                    f"but processor has {len(self.param_ranges)} parameters."
                )
            if not self._range_check(param_tensor):
                self.denormalize_param_dict(self.extract_param_dict(param_tensor))      # raises with the name
            scale, offset = self._affine(param_tensor.device)
            return self._packed_path[0](x, self.sample_rate, torch.addcmul(offset, param_tensor.to(torch.float32), scale))
        param_dict = self.extract_param_dict(param_tensor)
        denorm = self.denormalize_param_dict(param_dict, _checked=self._range_check(param_tensor))
        return self.process_fn(x, self.sample_rate, **denorm)

    def _affine(self, device):
        key = (str(device), tuple(self.param_ranges.values()))
        cache = self.__dict__.setdefault("_affine_cache", {})
        if key not in cache:
            lo = torch.tensor([r[0] for r in self.param_ranges.values()], dtype=torch.float32)
            hi = torch.tensor([r[1] for r in self.param_ranges.values()], dtype=torch.float32)
            cache.clear()
            cache[key] = ((hi - lo).to(device), lo.to(device))
        return cache[key]

    def process(self, x: torch.Tensor, *args):
        return self.process_fn(x, *args)

    def extract_param_dict(self, param_tensor: torch.Tensor):
        if param_tensor.shape[1] != len(self.param_ranges):
            raise ValueError(
                f"Parameter tensor has {param_tensor.shape[1]} parameters, "
                f"but processor has {len(self.param_ranges)} parameters."
            )
        return {name: param_tensor[:, i] for i, name in enumerate(self.param_ranges.keys())}

    @staticmethod
    def _range_check(param_tensor: torch.Tensor) -> bool:
        # one reduction + one host read for the whole tensor instead of 2 per parameter
        lo, hi = torch.aminmax(param_tensor.detach())
        bad = bool((lo < 0) | (hi > 1))
        return not bad

    def denormalize_param_dict(self, param_dict: dict, _checked=None):
        """(0, 1) -> physical ranges; raises ``ValueError`` on out-of-range input like the reference."""
        if _checked is None:
            _checked = all(self._range_check(v.reshape(1, -1)) for v in param_dict.values())
        if not _checked:
            for name, v in param_dict.items():
                if v.min() < 0 or v.max() > 1:
                    raise ValueError(f"Parameter {name} of is out of range.")
        out = {}
        for name, v in param_dict.items():
            lo, hi = self.param_ranges[name]
            out[name] = denormalize(v, hi, lo)
        return out


class Gain(Processor):
    def __init__(self, sample_rate: int, min_gain_db: float = -24.0, max_gain_db: float = 24.0):
        self.sample_rate = sample_rate
        self.process_fn = gain
        self.param_ranges = {"gain_db": (min_gain_db, max_gain_db)}


class Distortion(Processor):
    def __init__(self, sample_rate: int = 44100, min_gain_db: float = 0.0, max_gain_db: float = 24.0):
        self.sample_rate = sample_rate
        self.process_fn = distortion
        self.param_ranges = {"drive_db": (min_gain_db, max_gain_db)}


class ParametricEQ(Processor):
    def __init__(self, sample_rate: int, min_gain_db: float = -20.0, max_gain_db: float = 20.0,
                 min_q_factor: float = 0.1, max_q_factor: float = 6.0):
        self.sample_rate = sample_rate
        self.process_fn = parametric_eq
        self._packed_path = (_F.parametric_eq_packed, parametric_eq)
        g, q = (min_gain_db, max_gain_db), (min_q_factor, max_q_factor)
        top = (sample_rate // 2) - 1000
        cut = {"low_shelf": (20, 2000), "band0": (80, 2000), "band1": (2000, 8000), "band2": (8000, 12000),
               "band3": (12000, top), "high_shelf": (4000, top)}
        self.param_ranges = {}
        for sec, fr in cut.items():
            self.param_ranges[f"{sec}_gain_db"] = g
            self.param_ranges[f"{sec}_cutoff_freq"] = fr
            self.param_ranges[f"{sec}_q_factor"] = q


class _Dynamics(Processor):
    def __init__(self, sample_rate: int, min_threshold_db: float = -60.0, max_threshold_db: float = 0.0,
                 min_ratio: float = 1.0, max_ratio: float = 20.0, min_attack_ms: float = 5.0,
                 max_attack_ms: float = 100.0, min_release_ms: float = 5.0, max_release_ms: float = 100.0,
                 min_knee_db: float = 0.0, max_knee_db: float = 12.0, min_makeup_gain_db: float = 0.0,
                 max_makeup_gain_db: float = 12.0):
        self.sample_rate = sample_rate
        self.param_ranges = {
            "threshold_db": (min_threshold_db, max_threshold_db),
            "ratio": (min_ratio, max_ratio),
            "attack_ms": (min_attack_ms, max_attack_ms),
            "release_ms": (min_release_ms, max_release_ms),
            "knee_db": (min_knee_db, max_knee_db),
            "makeup_gain_db": (min_makeup_gain_db, max_makeup_gain_db),
        }


class Compressor(_Dynamics):
    def __init__(self, sample_rate: int, **kw):
        super().__init__(sample_rate, **kw)
        self.process_fn = compressor
        self._packed_path = (lambda x, sr, p: _F.dynamics_packed(0, x, sr, p), compressor)


class Expander(_Dynamics):
    """New: the reference advertises an expander but only stubs it (functional.py:402-403)."""

    def __init__(self, sample_rate: int, max_ratio: float = 4.0, **kw):
        super().__init__(sample_rate, max_ratio=max_ratio, **kw)
        self.process_fn = expander
        self._packed_path = (lambda x, sr, p: _F.dynamics_packed(1, x, sr, p), expander)


class NoiseShapedReverb(Processor):
    def __init__(self, sample_rate, min_band_gain: float = 0.0, max_band_gain: float = 1.0,
                 min_band_decay: float = 0.0, max_band_decay: float = 1.0, min_mix: float = 0.0,
                 max_mix: float = 1.0):
        self.sample_rate = sample_rate
        self.process_fn = noise_shaped_reverberation
        self._packed_path = (_F.noise_shaped_reverberation_packed, noise_shaped_reverberation)
        self.param_ranges = {f"band{i}_gain": (min_band_gain, max_band_gain) for i in range(12)}
        self.param_ranges.update({f"band{i}_decay": (min_band_decay, max_band_decay) for i in range(12)})
        self.param_ranges["mix"] = (min_mix, max_mix)
