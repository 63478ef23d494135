"""Which ingredient of a tensor-core mode costs how much accuracy: operand format x gate activations (tanh.approx / exact)
x 1x1 convs (TF32 tcgen05 / bf16-pair split tcgen05 / exact fp32), against the reference fixtures.  GPU box:  python tools/accuracy_matrix.py"""
import sys, itertools, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import load_golden
from test_oracle_vs_golden import build_from_meta
from oracle import dprnn_oracle as O
cases = sys.argv[1:] or ['spe_cat_r6_3s', 'tasnet_r6_3s', 'speech_att_r6']
for case in cases:
    meta, arr = load_golden(case)
    model = build_from_meta(meta).eval().cuda()
    mix, ref = torch.from_numpy(arr['mix']).cuda(), torch.from_numpy(arr['ref']).cuda()
    want = torch.from_numpy(arr['est'])
    for prec, fast, conv in itertools.product(('fp16', 'bf16'), (True, False), ('tf32', 'f32x2', 'fp32')):
        model.precision = prec
        model._engine.fast_act, model._engine.conv_kind = fast, conv
        with torch.no_grad():
            est = model(mix) if meta['cls'].endswith('DPRNNTasNet') else model(mix, ref, torch.tensor(float(meta['Tr'])))[0]
        e = est.cpu()
        sis = O.si_sdr_db(e.reshape(-1, e.shape[-1]), want.reshape(-1, want.shape[-1]))
        print(f'{case:18s} {prec} act={"approx" if fast else "exact "} conv={conv:5s}: '
              f'peak-normalised err {O.peak_rel_err(e, want):.2e}   SI-SDR {float(sis.min()):.1f} dB', flush=True)
