"""Pins the CPU oracle (oracle/dprnn_oracle.py) to fixtures produced by the live reference
(tests/golden/make_golden.py).  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, GOLDEN_CASES, load_golden, weight_fingerprint
from oracle import dprnn_oracle as O
import tss_with_dprnn_b200 as P

TOL = 2e-6   # fp32 re-association noise between two CPU evaluations of the same maths (peak-normalised)


def build_from_meta(meta):
    """Rebuild the seeded weights through OUR constructors (same RNG consumption as the reference's)."""
    cls = {'DPRNNTasNet': P.DPRNNTasNet, 'DPRNNSpeTasNet': P.DPRNNSpeTasNet,
           'DPRNNSpeIRATasNet': P.DPRNNSpeIRATasNet}[meta['cls'].rsplit('.', 1)[1]]
    torch.manual_seed(meta['wseed'])
    model = cls(**meta['kwargs'])
    if meta.get('lstm_wscale', 1.0) != 1.0:       # gates driven into saturation (tests/golden/make_speech_clips.py)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if '.rnn.weight_' in n:
                    p.mul_(meta['lstm_wscale'])
    model.train(meta['training'])
    return model


def oracle_cfg(meta):
    kw = meta['kwargs']
    return O.Config(**{k: kw[k] for k in ('input_size', 'feature_size', 'hidden_size', 'chunk_length', 'kernel_size',
                                          'hop_length', 'n_repeats', 'bidirectional', 'norm_type',
                                          'activation_type')},
                    fusion_type=kw.get('fusion_type', 'cat'))


def run_oracle(meta, model, mix, ref, fast=True):
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg = oracle_cfg(meta)
    name = meta['cls'].rsplit('.', 1)[1]
    stats = {}
    with torch.no_grad():
        if name == 'DPRNNTasNet':
            return {'est': O.tasnet_forward(mix, sd, cfg, fast=fast)}, stats
        fwd = O.ira_forward if name == 'DPRNNSpeIRATasNet' else O.spe_forward
        est, logits = fwd(mix, ref, torch.tensor(float(meta['Tr'])), sd, cfg, training=meta['training'],
                          new_stats=stats, fast=fast)
    return {'est': est, 'logits': logits}, stats


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_oracle_matches_reference_fixture(case):
    meta, arr = load_golden(case)
    torch.set_num_threads(os.cpu_count())
    model = build_from_meta(meta)
    fp = weight_fingerprint(model.state_dict())
    assert abs(fp - meta['weight_fingerprint']) <= 1e-9 * meta['weight_fingerprint'], \
        'seeded weights differ from the reference constructor (RNG consumption order changed?)'
    out, stats = run_oracle(meta, model, torch.from_numpy(arr['mix']), torch.from_numpy(arr['ref']))
    for k, v in out.items():
        assert v.shape == arr[k].shape
        assert O.peak_rel_err(v, torch.from_numpy(arr[k])) < TOL, k
    for k in arr:
        if k.startswith('stat:'):
            assert O.peak_rel_err(stats[k[5:]].float(), torch.from_numpy(arr[k])) < TOL, k


def test_explicit_lstm_equals_fused():
    meta, arr = load_golden('tasnet_r2')
    model = build_from_meta(meta)
    mix = torch.from_numpy(arr['mix'])[:1, :1500]
    a, _ = run_oracle(meta, model, mix, None, fast=True)
    b, _ = run_oracle(meta, model, mix, None, fast=False)
    assert O.peak_rel_err(a['est'], b['est']) < TOL


def test_index_maps_bit_exact():
    z = np.load(os.path.join(GOLDEN, 'index_maps.npz'))
    K, P_ = 250, 125
    for key in z.files:
        kind, L = key.split('_')
        L = int(L)
        if kind == 'unfold':
            assert np.array_equal(O.unfold_index_map(L, K, P_), z[key]), key
            assert z[key].shape[1] == O.n_chunks(L, K, P_)
        elif kind == 'foldcov':
            assert np.array_equal(O.fold_coverage(L, K, P_), z[key]), key
            assert (z[key] == 2).all()
        elif kind == 'nearest':
            assert np.array_equal(O.nearest_upsample_index((L - 2) // 2 + 1, L), z[key]), key


def test_fold_of_unfold_is_twice_identity():
    x = torch.randn(2, 8, 1234)
    y = O.overlap_add(O.segmentation(x, 250, 125), 1234, 250, 125)
    assert torch.equal(y, 2 * x)


@pytest.mark.skipif(not os.path.isdir('/root/reference/src/models'), reason='live reference not mounted')
def test_oracle_matches_live_reference_stagewise():
    """Per-stage check with forward hooks on the live reference modules (build container only)."""
    sys.path.insert(0, '/root/reference')
    from src.models.dprnn_spe import DPRNNSpeTasNet
    kw = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
              n_repeats=1, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0,
              fusion_type='att')
    torch.manual_seed(3)
    ref_model = DPRNNSpeTasNet(**kw).eval()
    sd = ref_model.state_dict()
    cfg = O.Config(n_repeats=1, fusion_type='att')
    g = torch.Generator().manual_seed(5)
    mix, aux = 0.05 * torch.randn(2, 3001, generator=g), 0.05 * torch.randn(2, 2801, generator=g)
    grabbed = {}
    hooks = [ref_model.separation.dprnn_blocks[0].register_forward_hook(lambda m, i, o: grabbed.update(blk_in=i[0], blk_out=o)),
             ref_model.separation.spk_encoder.register_forward_hook(lambda m, i, o: grabbed.update(spk=o)),
             ref_model.separation.bottleneck[1].register_forward_hook(lambda m, i, o: grabbed.update(fused=i[0]))]
    with torch.no_grad():
        ref_model(mix, aux, torch.tensor(2801.))
        enc = O.encoder(mix, sd['encoder.conv1d.weight'])
        e = O.speaker_embedding(O.encoder(aux, sd['encoder.conv1d.weight']), torch.tensor(2801.), sd, cfg)
        xn = O.chan_norm(enc, sd['separation.bottleneck.0.weight'], sd['separation.bottleneck.0.bias'], 1e-5)
        fused = O.fusion(e, xn, sd, cfg)
        assert O.peak_rel_err(fused, grabbed['fused']) < TOL
        assert O.peak_rel_err(e * (2800 // 27), grabbed['spk'].sum(-1)) < 1e-5
        blk = O.dprnn_block(grabbed['blk_in'], sd, 'separation.dprnn_blocks.0', cfg)
        assert O.peak_rel_err(blk, grabbed['blk_out']) < TOL
    for h in hooks:
        h.remove()


@pytest.mark.parametrize('L,K,P_', [(1337, 100, 50), (249, 250, 125), (1, 20, 10), (3999, 250, 125)])
def test_fold_can_apply_the_last_norm_and_residual(L, K, P_):
    """The identity behind dprnn_norm_residual_fold_prelu_h16: every chunk position (s, k) feeds exactly
    one output frame, so overlap_add(prelu(x + norm(y))) (dprnn.py:98-99,174,203-217) can be evaluated per frame t from the
    two positions s in [t // P + 1, min((t + K) // P, S - 1)], k = t + K - s P - the index arithmetic of the CUDA kernel,
    restated here with integer tensors and checked against the oracle's segmentation-map overlap-add."""
    torch.manual_seed(L)
    B, F = 2, 8
    S = O.n_chunks(L, K, P_)
    x, y = torch.randn(B, F, K, S, dtype=torch.float64), torch.randn(B, F, K, S, dtype=torch.float64)
    gamma, beta, a = torch.randn(F, dtype=torch.float64), torch.randn(F, dtype=torch.float64), 0.25
    v = x + O.chan_norm(y, gamma, beta, 1e-5)
    want = O.overlap_add(torch.where(v >= 0, v, a * v), L, K, P_)
    t = torch.arange(L)
    got = torch.zeros(B, F, L, dtype=torch.float64)
    covered = torch.zeros(L, dtype=torch.long)
    s_lo, s_hi = t // P_ + 1, torch.clamp((t + K) // P_, max=S - 1)
    for j in range(2):                                  # at most two chunks cover a frame when P = K / 2
        s = s_lo + j
        ok = s <= s_hi
        k = t + K - s * P_
        assert bool(((k >= 0) & (k < K))[ok].all())
        got[:, :, ok] += torch.where(v >= 0, v, a * v)[:, :, k[ok], s[ok]]
        covered += ok.long()
    assert bool((s_lo + 2 > s_hi).all())                # never a third chunk
    assert torch.equal(covered, torch.from_numpy(O.fold_coverage(L, K, P_)))
    assert torch.allclose(got, want, rtol=0, atol=1e-12)


@pytest.mark.skipif(not os.path.isdir('/root/reference/src/models'), reason='live reference not mounted')
@pytest.mark.parametrize('cls,T,Tr', [('tasnet', 24001, 0), ('spe', 30160, 24000), ('ira', 36720, 41000),
                                      ('spe', 111920, 107360), ('spe', 251, 400)])
def test_oracle_matches_live_reference_on_test_set_lengths(cls, T, Tr):
    """whole-model compare against the LIVE reference on the lengths its test set holds (median 36 720,
    maximum 111 920 samples; reference utterance longer or shorter than the mixture) - build container only, the fixtures
    cover the 3-s shape.  One block keeps the CPU forwards short; every stage's length arithmetic is exercised."""
    sys.path.insert(0, '/root/reference')
    from src.models.dprnn import DPRNNTasNet
    from src.models.dprnn_spe import DPRNNSpeTasNet
    from src.models.dprnn_spe_ira import DPRNNSpeIRATasNet
    kw = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
              n_repeats=1, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0)
    torch.manual_seed(T)
    g = torch.Generator().manual_seed(T + 1)
    B = 1 if T > 100000 else 2
    mix = 0.05 * torch.randn(B, T, generator=g)
    with torch.no_grad():
        if cls == 'tasnet':
            ref_model = DPRNNTasNet(**kw).eval()
            want = ref_model(mix)
            got = O.tasnet_forward(mix, ref_model.state_dict(), O.Config(n_repeats=1))
            assert want.shape == got.shape == (B, 2, T)
            assert O.peak_rel_err(got, want) < TOL
            return
        fusion = 'film' if cls == 'spe' else 'cat'
        ref_model = (DPRNNSpeTasNet if cls == 'spe' else DPRNNSpeIRATasNet)(**kw, fusion_type=fusion).eval()
        aux = 0.05 * torch.randn(B, Tr, generator=g)
        want, wl = ref_model(mix, aux, torch.tensor(float(Tr)))
        fwd = O.spe_forward if cls == 'spe' else O.ira_forward
        got, gl = fwd(mix, aux, torch.tensor(float(Tr)), ref_model.state_dict(), O.Config(n_repeats=1, fusion_type=fusion))
        assert want.shape == got.shape == (B, T)
        assert O.peak_rel_err(got, want) < TOL and O.peak_rel_err(gl, wl) < 1e-5
