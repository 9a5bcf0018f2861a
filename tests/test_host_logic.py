"""CPU tests of the host-side mirror of the reference interface: Processor parameter contract
(modules.py:21-231 upstream), signature/keyword compatibility, batch-sharding arithmetic."""
import inspect

import pytest
import torch

import dasp_pytorch_b200 as D
from dasp_pytorch_b200 import dist as ddist
from helpers import COMP_RANGES, eq_ranges


def test_signatures_match_reference_names():
    """keyword names are part of the ABI: Processor.process_normalized calls process_fn(x, sr, **named)"""
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(D.gain) == ["x", "sample_rate", "gain_db"]
    assert sig(D.distortion) == ["x", "sample_rate", "drive_db"]
    eq = ["x", "sample_rate"]
    for sec in ("low_shelf", "band0", "band1", "band2", "band3", "high_shelf"):
        eq += [f"{sec}_gain_db", f"{sec}_cutoff_freq", f"{sec}_q_factor"]
    assert sig(D.parametric_eq) == eq
    comp = ["x", "sample_rate", "threshold_db", "ratio", "attack_ms", "release_ms", "knee_db", "makeup_gain_db", "eps",
            "lookahead_samples"]
    assert sig(D.compressor) == comp and sig(D.expander) == comp
    p = inspect.signature(D.compressor).parameters
    assert p["eps"].default == 1e-8 and p["lookahead_samples"].default == 0
    rev = ["x", "sample_rate"] + [f"band{i}_gain" for i in range(12)] + [f"band{i}_decay" for i in range(12)] + [
        "mix", "num_samples", "num_bandpass_taps", "noise"]
    assert sig(D.noise_shaped_reverberation) == rev
    p = inspect.signature(D.noise_shaped_reverberation).parameters
    assert p["num_samples"].default == 65536 and p["num_bandpass_taps"].default == 1023
    assert p["noise"].kind is inspect.Parameter.KEYWORD_ONLY


def test_processor_ranges_and_order():
    eqp = D.ParametricEQ(44100)
    assert eqp.num_params == 18
    assert list(eqp.param_ranges.values()) == [tuple(r) for r in eq_ranges(44100)]
    assert list(eqp.param_ranges)[:3] == ["low_shelf_gain_db", "low_shelf_cutoff_freq", "low_shelf_q_factor"]
    cp = D.Compressor(44100)
    assert list(cp.param_ranges) == ["threshold_db", "ratio", "attack_ms", "release_ms", "knee_db", "makeup_gain_db"]
    assert list(cp.param_ranges.values()) == COMP_RANGES
    rv = D.NoiseShapedReverb(44100)
    assert rv.num_params == 25 and list(rv.param_ranges)[12] == "band0_decay" and list(rv.param_ranges)[24] == "mix"
    assert D.Gain(44100).param_ranges == {"gain_db": (-24.0, 24.0)}
    assert D.Distortion().param_ranges == {"drive_db": (0.0, 24.0)}


def test_reference_constructor_compat():
    """drop-in details of the reference classes (modules.py:94-121 upstream): ``Distortion(min, max)`` positional
    order, the reference's ``gain_db`` key still resolves, ``num_params`` is assignable like in every reference
    subclass (user-defined processors do ``self.num_params = len(self.param_ranges)``)."""
    d = D.Distortion(0.0, 12.0)
    assert d.param_ranges["drive_db"] == (0.0, 12.0) and d.param_ranges["gain_db"] == (0.0, 12.0)
    assert list(d.param_ranges) == ["drive_db"] and d.num_params == 1 and d.sample_rate == 44100
    assert D.Distortion(sample_rate=48000).sample_rate == 48000

    class Mine(D.Processor):
        def __init__(self):
            self.process_fn = lambda x, sr, **kw: x
            self.param_ranges = {"a": (0.0, 1.0), "b": (1.0, 2.0)}
            self.num_params = len(self.param_ranges)          # AttributeError in round 1

    m = Mine()
    assert m.num_params == 2
    assert m.process_normalized(torch.zeros(1, 1, 4), torch.rand(1, 2)).shape == (1, 1, 4)


def test_process_normalized_contract():
    """denormalisation, keyword dispatch and the reference's error behaviour (ValueError), with a CPU stub
    in place of the kernel so that no GPU is needed."""
    cp = D.Compressor(44100)
    seen = {}

    def stub(x, sample_rate, **kw):
        seen.update(kw, sample_rate=sample_rate)
        return x

    cp.process_fn = stub
    x = torch.zeros(3, 2, 8)
    p = torch.rand(3, 6)
    y = cp.process_normalized(x, p)
    assert y is x and seen["sample_rate"] == 44100
    for i, (name, (lo, hi)) in enumerate(cp.param_ranges.items()):
        assert torch.allclose(seen[name], p[:, i] * (hi - lo) + lo)
    with pytest.raises(ValueError):
        cp.process_normalized(x, torch.rand(3, 5))                    # wrong parameter count
    bad = p.clone()
    bad[1, 2] = 1.5
    with pytest.raises(ValueError, match="attack_ms"):
        cp.process_normalized(x, bad)                                 # out of (0, 1)
    bad[1, 2] = -0.1
    with pytest.raises(ValueError):
        cp.process_normalized(x, bad)
    # positional path (modules.py:53-54 upstream)
    got = []
    cp.process_fn = lambda x, *a: got.append(a) or x
    cp.process(x, 44100, *[p[:, i] for i in range(6)])
    assert got[0][0] == 44100 and len(got[0]) == 7


def test_shard_bounds_partition():
    for batch in (0, 1, 7, 8, 1024, 1023):
        for world in (1, 2, 3, 4, 8):
            spans = [ddist.shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert sizes == ddist.shard_sizes(batch, world)
    with pytest.raises(ValueError):
        ddist.shard_bounds(4, 2, 2)
    t = torch.arange(10)
    d = torch.arange(20)
    a, b = ddist.shard_tensors([t], 3, 1)[0], ddist.shard_tensors([d], 3, 1, rows_per_item=2)[0]
    assert a.tolist() == [4, 5, 6] and b.tolist() == [8, 9, 10, 11, 12, 13]
