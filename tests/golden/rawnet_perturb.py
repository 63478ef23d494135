"""Deterministic, name-keyed perturbation of the tensors whose DEFAULT initialisation would make a RawNet3 parity test
vacuous (BatchNorm running statistics are 0/1, AFMS alpha is 1, InstanceNorm affine is 1/0).  Applied identically to the
reference model (fixture generation) and to the model under test, so the fixtures need not store 16 M weights."""
import zlib

import torch


def perturb_rawnet_state(model):
    sd = model.state_dict()
    new = {}
    for k, v in sd.items():
        if 'spk_encoder' not in k or not v.dtype.is_floating_point:
            continue
        g = torch.Generator().manual_seed(zlib.crc32(k.encode()))
        if k.endswith('running_mean'):
            new[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('running_var'):
            new[k] = 0.6 + 0.8 * torch.rand(v.shape, generator=g)
        elif k.endswith('afms.alpha'):
            new[k] = 1.0 + 0.2 * torch.randn(v.shape, generator=g)
        elif 'preprocess.1.' in k or ('.bn' in k and (k.endswith('.weight') or k.endswith('.bias'))) \
                or ('attention.2.' in k and (k.endswith('.weight') or k.endswith('.bias'))):
            base = 1.0 if k.endswith('weight') else 0.0
            new[k] = base + 0.1 * torch.randn(v.shape, generator=g)
    missing = model.load_state_dict({**sd, **new}, strict=True)
    return len(new)
