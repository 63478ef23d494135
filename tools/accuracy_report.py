"""Accuracy of every precision mode of the CUDA path against the fixtures produced by the live reference
(tests/golden/*.npz).  Run on the GPU box:  python tools/accuracy_report.py > profiles/rN_accuracy_report.txt"""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import load_golden
from test_oracle_vs_golden import build_from_meta
from oracle import dprnn_oracle as O
MODES = (('fp32', None), ('fp16', True), ('fp16', False), ('bf16', True), ('bf16', False))
for case in ('spe_cat_r6_3s', 'tasnet_r6_3s', 'speech_att_r6', 'speech_cat_r6_wx3', 'spe_att_r2_eval', 'ira_cat_r2_eval'):
    meta, arr = load_golden(case)
    model = build_from_meta(meta).eval().cuda()
    mix, ref = torch.from_numpy(arr['mix']).cuda(), torch.from_numpy(arr['ref']).cuda()
    want = torch.from_numpy(arr['est'])
    for prec, res16 in MODES:
        model.precision = prec
        if res16 is not None:
            model._engine.residual_bf16 = res16
        with torch.no_grad():
            est = model(mix) if meta['cls'].endswith('DPRNNTasNet') else model(mix, ref, torch.tensor(float(meta['Tr'])))[0]
        e = est.cpu()
        sis = O.si_sdr_db(e.reshape(-1, e.shape[-1]), want.reshape(-1, want.shape[-1]))
        label = prec if res16 is None else f"{prec}+{'16-bit' if res16 else 'fp32'} residual"
        print(f'{case:18s} {label:22s}: peak-normalised err {O.peak_rel_err(e, want):.2e}   '
              f'SI-SDR(ours, reference fp32) min {float(sis.min()):.1f} dB', flush=True)
