"""Golden fixtures for DPRNN-RawNet (cfg 4) from the LIVE reference classes.

    python tests/golden/make_golden_rawnet.py        # build container only (needs /root/reference)

The reference's RawNet3 imports ``asteroid_filterbanks`` (third-party, absent: SURVEY.md section 8c).  It is stubbed
here with the restatement in oracle/rawnet_oracle.py - so these fixtures pin everything the reference itself contains
(RawNet3.forward, Bottle2neck, AFMS, PreEmphasis, DPRNNRawNet / DPRNNRawNetTasNet and the module tree / state_dict
layout) and leave exactly the sinc filterbank arithmetic unpinned.  Weights: the reference's default initialisation
under torch.manual_seed(0), then tests/golden/rawnet_perturb.py."""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference')

from oracle import rawnet_oracle as RO  # noqa: E402
from rawnet_perturb import perturb_rawnet_state  # noqa: E402

stub = types.ModuleType('asteroid_filterbanks')
stub.Encoder, stub.ParamSincFB = RO.Encoder, RO.ParamSincFB
sys.modules['asteroid_filterbanks'] = stub

from src.models.dprnn_rawnet import DPRNNRawNetTasNet  # noqa: E402

KW = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
          n_repeats=1, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0,
          embeddings_size=256, fusion_type='att')


def fingerprint(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.dtype.is_floating_point))


def main():
    torch.manual_seed(0)
    model = DPRNNRawNetTasNet(**KW).eval()
    fp0 = fingerprint(model.state_dict())
    n = perturb_rawnet_state(model)
    B, T, Tr = 2, 6000, 12000
    g = torch.Generator().manual_seed(77)
    mix = 0.05 * torch.randn(B, T, generator=g)
    ref = 0.05 * torch.randn(B, Tr, generator=g)
    with torch.no_grad():
        emb = model.separation.spk_encoder(ref)
        est, logits = model(mix, ref)
    sd = model.state_dict()
    layout = {k: [list(v.shape), str(v.dtype)] for k, v in sd.items()}
    meta = dict(cls='src.models.dprnn_rawnet.DPRNNRawNetTasNet', kwargs=KW, B=B, T=T, Tr=Tr, wseed=0, iseed=77,
                weight_fingerprint_seeded=fp0, weight_fingerprint=fingerprint(sd), perturbed=n, training=False)
    np.savez_compressed(os.path.join(HERE, 'rawnet_att_r1_eval.npz'), meta=json.dumps(meta), mix=mix.numpy(),
                        ref=ref.numpy(), emb=emb.numpy(), est=est.numpy(), logits=logits.numpy())
    json.dump(layout, open(os.path.join(HERE, 'state_dict_layout_rawnet.json'), 'w'), indent=0)
    print('rawnet_att_r1_eval:', {k: tuple(v.shape) for k, v in dict(emb=emb, est=est, logits=logits).items()},
          'entries', len(layout), 'params', sum(p.numel() for p in model.parameters()))


if __name__ == '__main__':
    main()
