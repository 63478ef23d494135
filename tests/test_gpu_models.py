"""Whole-model parity of the CUDA path (through the model classes -> C ABI) against (a) fixtures
produced by the live reference and (b) the CPU oracle, plus size-independent properties at the
BASELINE shapes.  Tolerance of the fp32 mode: peak-normalised max error <= 1e-3 (north_star); the
exact-fp32 kernels are held to a much tighter 2e-5 here."""
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden, weight_fingerprint
from oracle import dprnn_oracle as O
import tss_with_dprnn_b200 as P
from test_oracle_vs_golden import build_from_meta, run_oracle

pytestmark = pytest.mark.gpu
TOL_FP32 = 2e-5
KW = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
          n_repeats=6, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0)


def run_cuda(meta, model, mix, ref):
    model = model.cuda()
    with torch.no_grad():
        if meta['cls'].endswith('DPRNNTasNet'):
            return {'est': model(mix.cuda()).cpu()}
        est, logits = model(mix.cuda(), ref.cuda(), torch.tensor(float(meta['Tr'])))
    return {'est': est.cpu(), 'logits': logits.cpu()}


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_cuda_matches_reference_fixture(case):
    meta, arr = load_golden(case)
    model = build_from_meta(meta)
    assert abs(weight_fingerprint(model.state_dict()) - meta['weight_fingerprint']) <= 1e-9 * meta['weight_fingerprint']
    out = run_cuda(meta, model, torch.from_numpy(arr['mix']), torch.from_numpy(arr['ref']))
    for k, v in out.items():
        assert v.shape == arr[k].shape, k
        err = O.peak_rel_err(v, torch.from_numpy(arr[k]))
        assert err < TOL_FP32, (k, err)
    sd = model.state_dict()
    for k in arr:
        if k.startswith('stat:'):        # BatchNorm running statistics after a train-mode forward
            assert O.peak_rel_err(sd[k[5:]].cpu().float(), torch.from_numpy(arr[k])) < TOL_FP32, k
    if meta['training'] and not meta['cls'].endswith('DPRNNTasNet'):
        n = 2 if 'IRA' in meta['cls'] else 1
        assert int(sd['separation.spk_encoder.2.batch_norm1.num_batches_tracked']) == n


@pytest.mark.parametrize('T,Tr', [(24001, 16000), (30160, 24000), (377, 1000), (251, 300)])
def test_cuda_matches_oracle_odd_lengths(T, Tr):
    """Lengths the fixtures do not cover: odd T, ref length != mix length, the shortest legal input."""
    torch.manual_seed(4)
    model = P.DPRNNSpeTasNet(**dict(KW, n_repeats=1), fusion_type='att').eval()
    meta = dict(cls='x.DPRNNSpeTasNet', kwargs=dict(KW, n_repeats=1, fusion_type='att'), Tr=Tr, training=False)
    g = torch.Generator().manual_seed(T)
    mix, ref = 0.05 * torch.randn(2, T, generator=g), 0.05 * torch.randn(2, Tr, generator=g)
    want, _ = run_oracle(meta, model, mix, ref)
    got = run_cuda(meta, model, mix, ref)
    for k in want:
        assert O.peak_rel_err(got[k], want[k]) < TOL_FP32, k


def test_embedding_injection_matches_oracle():
    """cfg-4 style: RawNet-sized embedding (E=256) injected on both sides, attention fusion."""
    kw = dict(KW, n_repeats=1, fusion_type='att', embeddings_size=256)
    torch.manual_seed(5)
    model = P.DPRNNSpeTasNet(**kw).eval()
    g = torch.Generator().manual_seed(99)
    mix, emb = 0.05 * torch.randn(2, 6000, generator=g), torch.randn(2, 256, generator=g)
    cfg = O.Config(n_repeats=1, fusion_type='att', embeddings_size=256)
    with torch.no_grad():
        want, wl = O.spe_forward(mix, None, None, {k: v.clone() for k, v in model.state_dict().items()}, cfg, embedding=emb)
        model = model.cuda()
        got, gl = model.forward_with_embedding(mix.cuda(), emb.cuda())
    assert O.peak_rel_err(got.cpu(), want) < TOL_FP32
    assert O.peak_rel_err(gl.cpu(), wl) < TOL_FP32


def test_batch_independence_bit_exact():
    """Utterances are independent: a batched forward equals per-utterance forwards bit for bit,
    which is also what makes sharding the batch over GPUs exact."""
    torch.manual_seed(6)
    model = P.DPRNNSpeTasNet(**dict(KW, n_repeats=2), fusion_type='cat').eval().cuda()
    g = torch.Generator().manual_seed(8)
    mix, ref = (0.05 * torch.randn(5, 8000, generator=g)).cuda(), (0.05 * torch.randn(5, 8000, generator=g)).cuda()
    rl = torch.tensor(8000.)
    with torch.no_grad():
        est, logits = model(mix, ref, rl)
        for b in range(5):
            e1, l1 = model(mix[b:b + 1], ref[b:b + 1], rl)
            assert torch.equal(e1[0], est[b]) and torch.equal(l1[0], logits[b]), b
        est2, _ = model(mix, ref, rl)
    assert torch.equal(est, est2)            # run-to-run determinism


def test_full_size_properties_cfg2():
    """cfg 2 shape (3 s, full depth) at a batch the fp32 path finishes quickly: finite, bounded by the mixture
    encoder gain (sigmoid masks in [0,1] => |est| cannot exceed what mask=1 gives), deterministic."""
    torch.manual_seed(0)
    model = P.DPRNNSpeTasNet(**KW, fusion_type='cat').eval().cuda()
    g = torch.Generator().manual_seed(1234)
    mix, ref = (0.05 * torch.randn(8, 24000, generator=g)).cuda(), (0.05 * torch.randn(8, 24000, generator=g)).cuda()
    with torch.no_grad():
        est, logits = model(mix, ref, torch.tensor(24000.))
    assert est.shape == (8, 24000) and logits.shape == (8, 251)
    assert torch.isfinite(est).all() and torch.isfinite(logits).all()
    w_enc = model.encoder.conv1d.weight.detach().abs().sum()
    w_dec = model.decoder.weight.detach().abs().max()
    assert est.abs().max() <= 2 * mix.abs().max() * w_enc * w_dec


def test_rejects_cpu_tensors_and_grad_training():
    model = P.DPRNNSpeTasNet(**dict(KW, n_repeats=1)).cuda()
    with pytest.raises(RuntimeError, match='no CPU path'):
        with torch.no_grad():
            model(torch.zeros(1, 4000), torch.zeros(1, 4000), torch.tensor(4000.))
    # train() + autograd: every model class returns tensors attached to the hand-written backward (train.py)
    model.train()
    est, logits = model(0.1 * torch.randn(2, 4000).cuda(), 0.1 * torch.randn(2, 4000).cuda(), torch.tensor(4000.))
    assert est.requires_grad and logits.requires_grad
    ira = P.DPRNNSpeIRATasNet(**dict(KW, n_repeats=1)).cuda().train()
    est, logits = ira(0.1 * torch.randn(2, 4000).cuda(), 0.1 * torch.randn(2, 4000).cuda(), torch.tensor(4000.))
    assert est.requires_grad and logits.requires_grad


# ---------------------------------------------------------------------------------------------------------
# bf16 tensor-core mode (north_star: "test-set SI-SDR within 0.05 dB in the bf16-GEMM mode").  The test set and the
# checkpoints are not available, so the proxy of SURVEY.md section 8c(v) is used: seeded weights, synthetic mixtures,
# (a) SI-SDR of our bf16 output against the reference fp32 output, (b) the change of SI-SDR towards a synthetic target
# placed so that the fp32 output scores ~13 dB (the published operating point).
# ---------------------------------------------------------------------------------------------------------
def _bf16_vs_fp32(case, fast_act=True, precision='bf16', residual16=None):
    meta, arr = load_golden(case)
    model = build_from_meta(meta).cuda()
    model.precision = precision
    if residual16 is not None:
        model._engine.residual_bf16 = residual16
    model._engine.fast_act = fast_act
    mix, ref = torch.from_numpy(arr['mix']).cuda(), torch.from_numpy(arr['ref']).cuda()
    with torch.no_grad():
        if meta['cls'].endswith('DPRNNTasNet'):
            est = model(mix).cpu().reshape(-1, arr['est'].shape[-1])
        else:
            est = model(mix, ref, torch.tensor(float(meta['Tr'])))[0].cpu()
    want = torch.from_numpy(arr['est']).reshape(est.shape)
    return est, want


@pytest.mark.parametrize('case', ['spe_cat_r6_3s', 'tasnet_r6_3s', 'spe_att_r2_eval', 'ira_cat_r2_eval', 'spe_cat_uni_r2',
                                  'speech_att_r6', 'speech_cat_r6_wx3'])
@pytest.mark.parametrize('fast_act', [True, False])
def test_bf16_mode_sisdr(case, fast_act):
    est, want = _bf16_vs_fp32(case, fast_act)
    assert torch.isfinite(est).all()
    sdr_vs_ref = O.si_sdr_db(est, want)
    if case in SATURATED:       # gates in saturation: emulated 39.5 dB / 1.4e-2 (tools/emulate_precision.py --wscale 3)
        assert sdr_vs_ref.min() > 33.0, sdr_vs_ref
        assert O.peak_rel_err(est, want) < 4e-2
        return
    # measured (profiles/r2_accuracy_report.txt): 51-53 dB / 3.2e-3 on the full-depth synthetic fixtures, 48.1 dB / 7.1e-3
    # on the real-speech one, 67-68 dB / 8e-4 on the 2-block ones; thresholds = measured - 5 dB / x1.7 (a regression of
    # the operand handling shows up as >= 6 dB)
    floor_db, ceil_err = (43.0, 1.2e-2) if case.startswith('speech') else (46.0, 6e-3) if '_r6' in case else (60.0, 2e-3)
    assert sdr_vs_ref.min() > floor_db, sdr_vs_ref
    assert O.peak_rel_err(est, want) < ceil_err
    g = torch.Generator().manual_seed(77)
    noise = torch.randn(want.shape, generator=g)
    noise = noise * (want.pow(2).sum(-1, keepdim=True) / noise.pow(2).sum(-1, keepdim=True) / 10 ** 1.3).sqrt()
    target = want + noise                                       # SI-SDR(want, target) ~ 13 dB
    delta = (O.si_sdr_db(est, target) - O.si_sdr_db(want, target)).abs()
    assert delta.max() < 0.05, delta


SATURATED = {'speech_cat_r6_wx3'}
TOL_NORTH_STAR = 1e-3     # north_star: estimated sources within a max relative (peak-normalised) error of 1e-3


@pytest.mark.parametrize('case', GOLDEN_CASES)
@pytest.mark.parametrize('fast_act', [True, False])
def test_fp16_mode_meets_1e3(case, fast_act):
    """The tensor-core mode that meets north_star's floating-point tolerance: fp16 operands (11 significand bits) on the
    same tcgen05 kernels as the bf16 mode, 16-bit residual stream included.  Every reference fixture (both full-depth
    3-s ones, every 2-block one, train-mode BatchNorm ones) must be within 1e-3 peak-normalised."""
    est, want = _bf16_vs_fp32(case, fast_act, precision='fp16')
    assert torch.isfinite(est).all()
    err = O.peak_rel_err(est, want)
    if case in SATURATED:
        # LSTM weights x3 (gates in saturation): the network amplifies ANY perturbation ~6x more than at its default
        # init - TF32 1x1 convs alone give 1.4e-3, tanh.approx alone 9e-4 (tools/emulate_precision.py --wscale 3) - so
        # only the exact fp32 mode holds 1e-3 there (test_cuda_matches_reference_fixture); fp16 is held to its emulated
        # 2.6e-3 + margin, bf16 lands at 1.4e-2
        assert err < 5e-3, err
        assert O.si_sdr_db(est, want).min() > 50.0
        return
    assert err < TOL_NORTH_STAR, err
    assert O.si_sdr_db(est, want).min() > 60.0


@pytest.mark.parametrize('case', ['spe_cat_r6_3s', 'tasnet_r6_3s'])
def test_fp16_mode_fp32_master(case):
    """fp16 operands with the fp32 master copy of the residual stream (Engine.residual_bf16 = False)."""
    est, want = _bf16_vs_fp32(case, True, precision='fp16', residual16=False)
    assert O.peak_rel_err(est, want) < TOL_NORTH_STAR


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
def test_cfg2_batch64_streams_graph_equals_b1(precision):
    """Parity AT the benchmarked configuration: B = 64 x 3 s, full depth, 3 utterance groups on concurrent streams,
    CUDA-graph replay - every utterance must equal its own B = 1 forward bit for bit, and utterance 0 (the fixture's
    mixture / reference) must agree with the reference fixture as the B = 1 tests require."""
    meta, arr = load_golden('spe_cat_r6_3s')
    model = build_from_meta(meta).eval().cuda()
    model.precision = precision
    g = torch.Generator().manual_seed(64)
    mix, ref = 0.05 * torch.randn(64, 24000, generator=g), 0.05 * torch.randn(64, 24000, generator=g)
    mix[0], ref[0] = torch.from_numpy(arr['mix'][0]), torch.from_numpy(arr['ref'][0])
    mix, ref, rl = mix.cuda(), ref.cuda(), torch.tensor(24000.)
    with torch.no_grad():
        model.n_streams = 3
        outs = [model(mix, ref, rl) for _ in range(3)]        # eager, capture, replay
        torch.cuda.synchronize()
        for e, l in outs[1:]:
            assert torch.equal(e, outs[0][0]) and torch.equal(l, outs[0][1])
        est, logits = outs[2]
        model.n_streams = 1
        for b in (0, 1, 21, 22, 42, 63):                      # first / last utterances of every stream group
            e1, l1 = model(mix[b:b + 1], ref[b:b + 1], rl)
            assert torch.equal(e1[0], est[b]) and torch.equal(l1[0], logits[b]), b
    err = O.peak_rel_err(est[:1].cpu(), torch.from_numpy(arr['est']))
    assert err < (TOL_NORTH_STAR if precision == 'fp16' else 6e-3), err


@pytest.mark.parametrize('precision', ['fp32', 'bf16', 'fp16'])
def test_stream_groups_bit_exact(precision):
    """Splitting the batch over concurrent streams inside forward must not change a single bit."""
    torch.manual_seed(6)
    model = P.DPRNNSpeTasNet(**dict(KW, n_repeats=2), fusion_type='film').eval().cuda()
    model.precision = precision
    g = torch.Generator().manual_seed(9)
    mix, ref = (0.05 * torch.randn(7, 6000, generator=g)).cuda(), (0.05 * torch.randn(7, 6000, generator=g)).cuda()
    with torch.no_grad():
        model.n_streams = 1
        e1, l1 = model(mix, ref, torch.tensor(6000.))
        model.n_streams = 3
        e3, l3 = model(mix, ref, torch.tensor(6000.))
        torch.cuda.synchronize()
    assert torch.equal(e1, e3) and torch.equal(l1, l3)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
@pytest.mark.parametrize('cls', ['spe', 'ira', 'bss'])
def test_fused_input_norm_whole_model_bit_identical(precision, cls):
    """Engine.fuse_norm (default): the half-block's norm + residual moved into the next LSTM kernel changes no bit."""
    torch.manual_seed(9)
    kw = dict(KW, n_repeats=2)
    model = {'spe': lambda: P.DPRNNSpeTasNet(**kw, fusion_type='film'), 'ira': lambda: P.DPRNNSpeIRATasNet(**kw, fusion_type='cat'),
             'bss': lambda: P.DPRNNTasNet(**kw)}[cls]().eval().cuda()
    model.precision = precision
    g = torch.Generator().manual_seed(10)
    mix, ref = (0.05 * torch.randn(3, 5000, generator=g)).cuda(), (0.05 * torch.randn(3, 4000, generator=g)).cuda()
    args = (mix,) if cls == 'bss' else (mix, ref, torch.tensor(4000.))
    outs = []
    with torch.no_grad():
        for fuse in (True, False):
            model._engine.fuse_norm = fuse
            n0 = P.lib().launches
            o = model(*args)
            outs.append(((o,) if cls == 'bss' else o, P.lib().launches - n0))
    (a, na), (b, nb) = outs
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert nb - na == (3 if cls != 'ira' else 6)          # one norm launch less per half-block but the last (2 blocks)


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
@pytest.mark.parametrize('cls', ['spe', 'ira', 'bss'])
def test_fold_fused_whole_model_bit_identical(precision, cls):
    """Engine.fold_fused (default): the last half-block's norm + residual applied by the fold, no fp32 [B,S,K,F] tensor -
    no bit changes against the separate kernels, two launches become one per masker pass."""
    torch.manual_seed(11)
    kw = dict(KW, n_repeats=2)
    model = {'spe': lambda: P.DPRNNSpeTasNet(**kw, fusion_type='att'), 'ira': lambda: P.DPRNNSpeIRATasNet(**kw, fusion_type='cat'),
             'bss': lambda: P.DPRNNTasNet(**kw)}[cls]().eval().cuda()
    model.precision = precision
    g = torch.Generator().manual_seed(12)
    mix, ref = (0.05 * torch.randn(3, 5001, generator=g)).cuda(), (0.05 * torch.randn(3, 4000, generator=g)).cuda()
    args = (mix,) if cls == 'bss' else (mix, ref, torch.tensor(4000.))
    outs = []
    with torch.no_grad():
        for fused in (True, False):
            model._engine.fold_fused = fused
            n0 = P.lib().launches
            o = model(*args)
            outs.append(((o,) if cls == 'bss' else o, P.lib().launches - n0))
    (a, na), (b, nb) = outs
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert nb - na == (1 if cls != 'ira' else 2)
