"""TEST INFRASTRUCTURE: weights and inputs that both sides of a parity test can rebuild from a seed.

The host rows of the hot path (reference models/encoder.py, models/fusion.py, models/model*.py) carry tens of
millions of parameters (ResNet-50 trunk, 256-channel ConvGRU); their state_dicts cannot travel as fixtures.  Instead
the golden generator (``oracle/make_golden.py``, run where /root/reference exists) and the GPU tests both fill a
``state_dict`` with ``seeded_state_dict``: one generator per key, seeded from the key's NAME, so the values do not
depend on construction order, on which optional sub-modules exist, or on torch's module-init RNG consumption.  Only the
inputs' seeds and the reference outputs are stored under tests/golden/.
"""
import zlib

import torch


def _gen(seed, key):
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))


GAINS = (('feature_extraction', 0.8), ('', 1.4))     # first substring match wins; the 16-block residual trunk needs < 1


def seeded_state_dict(module, seed):
    """A full state_dict for ``module`` (parameters and buffers) with deterministic, well-conditioned values:
    conv / linear weights ~ N(0, gain^2 / fan_in) (GAINS keeps activations O(1) through the ReLU / residual stacks),
    norm scales in [0.6, 1.4], biases and running means ~ N(0, 0.1^2), running variances in [0.6, 1.4]."""
    out = {}
    for key, ref in module.state_dict().items():
        g = _gen(seed, key)
        leaf = key.rsplit('.', 1)[-1]
        if leaf == 'num_batches_tracked' or not ref.is_floating_point():
            out[key] = ref.clone()
        elif leaf == 'running_var':
            out[key] = torch.rand(ref.shape, generator=g) * 0.8 + 0.6
        elif leaf == 'running_mean' or leaf == 'bias' or ref.dim() == 0:
            out[key] = torch.randn(ref.shape, generator=g) * 0.1
        elif ref.dim() == 1:                        # norm scale (BatchNorm / LayerNorm weight) or a 1-D table
            out[key] = torch.rand(ref.shape, generator=g) * 0.8 + 0.6
        else:
            owner = module.get_submodule(key.rsplit('.', 1)[0]) if '.' in key else module
            if type(owner).__name__.startswith('ConvTranspose'):      # weight [cin, cout, *k]: taps that hit one output
                taps = ref[0, 0].numel()
                for st in owner.stride:
                    taps //= st
                fan_in = ref.shape[0] * max(taps, 1)
            else:
                fan_in = ref[0].numel() if ref.dim() > 1 else ref.numel()
            gain = next(gn for sub, gn in GAINS if sub in key)
            out[key] = torch.randn(ref.shape, generator=g) * (gain / max(fan_in, 1) ** 0.5)
        out[key] = out[key].to(ref.dtype)
    return out


def seeded_tensor(seed, name, shape, kind='randn', scale=1.0):
    g = _gen(seed, name)
    t = torch.randn(shape, generator=g) if kind == 'randn' else torch.rand(shape, generator=g)
    return t * scale


def load_seeded(module, seed, strict=True, only_prefixes=None):
    """Fill ``module`` in place.  ``only_prefixes``: restrict to keys of a LARGER reference module (e.g. the mirror has no
    pose networks): the values of a key depend on its name only, so ``load_seeded(mirror, s)`` and
    ``load_seeded(reference, s)`` agree on every shared key."""
    sd = seeded_state_dict(module, seed)
    if only_prefixes is not None:
        sd = {k: v for k, v in sd.items() if k.startswith(tuple(only_prefixes))}
    module.load_state_dict(sd, strict=strict and only_prefixes is None)
    return module


def subsample(t, step=4):
    """every step-th pixel of the two trailing dims (keeps image fixtures small)"""
    return t[..., ::step, ::step].contiguous()
