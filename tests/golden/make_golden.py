"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors and its checkpoints are missing, so
the pins are: the reference's own modules (``src/models``) built with their default initialisation
under ``torch.manual_seed(seed)`` and run on seeded synthetic mixtures.  Each ``*.npz`` stores the
constructor kwargs, the seeds, the inputs, the reference outputs and a fingerprint of the seeded
weights (so a consumer that rebuilds the weights from the seed can tell a RNG mismatch from a kernel
bug).  ``index_maps.npz`` stores integer maps produced by the very torch ops the reference calls
(F.unfold / F.fold / nn.Upsample) for bit-exact checks.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))

BASE = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2,
            hop_length=125, n_repeats=2, bidirectional=True, norm_type='ln',
            activation_type='sigmoid', dropout=0)


def fingerprint(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.dtype.is_floating_point))


def inputs(B, T, Tr, seed):
    g = torch.Generator().manual_seed(seed)
    mix = 0.05 * torch.randn(B, T, generator=g)
    ref = 0.05 * torch.randn(B, Tr, generator=g)
    return mix, ref


def run_case(name, cls_path, kwargs, B, T, Tr, training, wseed=0, iseed=1234):
    mod, cls = cls_path.rsplit('.', 1)
    cls = getattr(__import__(mod, fromlist=[cls]), cls)
    torch.manual_seed(wseed)
    model = cls(**kwargs)
    model.train(training)
    fp = fingerprint(model.state_dict())
    mix, ref = inputs(B, T, Tr, iseed)
    out = {}
    with torch.no_grad():
        if cls.__name__ == 'DPRNNTasNet':
            out['est'] = model(mix).numpy()
        else:
            est, logits = model(mix, ref, torch.tensor(float(Tr)))
            out['est'], out['logits'] = est.numpy(), logits.numpy()
            if training:   # BatchNorm running statistics after the forward (InferencerSpe path, Appendix B)
                sd = model.state_dict()
                for k in sd:
                    if 'running_' in k:
                        out['stat:' + k] = sd[k].numpy()
    meta = dict(cls=cls_path, kwargs=kwargs, B=B, T=T, Tr=Tr, training=training, wseed=wseed,
                iseed=iseed, weight_fingerprint=fp, torch=torch.__version__)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), meta=json.dumps(meta), mix=mix.numpy(),
                        ref=ref.numpy(), **out)
    print(name, {k: v.shape for k, v in out.items() if not k.startswith('stat:')}, 'fp', fp)


def index_maps():
    from torch.nn.functional import unfold, fold
    K, P = 250, 125
    out = {}
    for L in (23999, 3999, 24000, 30159, 249, 1):
        x = torch.arange(1, L + 1, dtype=torch.float64).view(1, 1, L, 1)       # frame t stored as t+1, 0 = padding
        u = unfold(x, kernel_size=(K, 1), padding=(K, 0), stride=(P, 1))      # [1, K, S]
        out[f'unfold_{L}'] = (u[0].long() - 1).numpy()
        cov = fold(torch.ones_like(u), (L, 1), kernel_size=(K, 1), padding=(K, 0), stride=(P, 1))
        out[f'foldcov_{L}'] = cov.reshape(-1).long().numpy()
    for L in (23999, 3999, 24000, 30159, 111919, 5, 4):
        La = (L - 2) // 2 + 1
        src = torch.nn.Upsample(size=L, mode='nearest')(torch.arange(La, dtype=torch.float32).view(1, 1, La))
        out[f'nearest_{L}'] = src.reshape(-1).long().numpy()
    np.savez_compressed(os.path.join(HERE, 'index_maps.npz'), **out)
    print('index_maps', len(out))


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count())
    index_maps()
    run_case('tasnet_r2', 'src.models.dprnn.DPRNNTasNet', BASE, 2, 4000, 0, False)
    for ft in ('cat', 'add', 'mul', 'film', 'att'):
        run_case(f'spe_{ft}_r2_eval', 'src.models.dprnn_spe.DPRNNSpeTasNet', dict(BASE, fusion_type=ft),
                 2, 4000, 4300, False)
    run_case('spe_cat_r2_train', 'src.models.dprnn_spe.DPRNNSpeTasNet', dict(BASE, fusion_type='cat'),
             2, 4000, 4300, True)
    run_case('spe_att_r2_train_b1', 'src.models.dprnn_spe.DPRNNSpeTasNet', dict(BASE, fusion_type='att'),
             1, 4001, 3777, True)
    run_case('spe_film_gln_relu_r2', 'src.models.dprnn_spe.DPRNNSpeTasNet',
             dict(BASE, fusion_type='film', norm_type='gLN', activation_type='relu'), 2, 4000, 4300, False)
    run_case('spe_cat_uni_r2', 'src.models.dprnn_spe.DPRNNSpeTasNet',
             dict(BASE, fusion_type='cat', bidirectional=False), 2, 4000, 4300, False)
    run_case('ira_cat_r2_eval', 'src.models.dprnn_spe_ira.DPRNNSpeIRATasNet', dict(BASE, fusion_type='cat'),
             2, 4000, 4300, False)
    run_case('ira_cat_r2_train', 'src.models.dprnn_spe_ira.DPRNNSpeIRATasNet', dict(BASE, fusion_type='cat'),
             1, 4000, 4300, True)
    # the headline shape at full depth (cfg 2 with B=1) and cfg 1
    run_case('spe_cat_r6_3s', 'src.models.dprnn_spe.DPRNNSpeTasNet', dict(BASE, fusion_type='cat', n_repeats=6),
             1, 24000, 24000, False)
    run_case('tasnet_r6_3s', 'src.models.dprnn.DPRNNTasNet', dict(BASE, n_repeats=6), 1, 24000, 0, False)
