import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import load_golden
from test_oracle_vs_golden import build_from_meta
from oracle import dprnn_oracle as O
for case in ('spe_cat_r6_3s', 'tasnet_r6_3s', 'spe_att_r2_eval', 'ira_cat_r2_eval'):
    meta, arr = load_golden(case)
    model = build_from_meta(meta).eval().cuda()
    mix, ref = torch.from_numpy(arr['mix']).cuda(), torch.from_numpy(arr['ref']).cuda()
    want = torch.from_numpy(arr['est'])
    for prec in ('fp32', 'bf16', 'bf16+residual_bf16'):
        model.precision = prec.split('+')[0]
        model._engine.residual_bf16 = prec.endswith('residual_bf16')
        with torch.no_grad():
            est = model(mix) if meta['cls'].endswith('DPRNNTasNet') else model(mix, ref, torch.tensor(float(meta['Tr'])))[0]
        e = est.cpu()
        sis = O.si_sdr_db(e.reshape(-1, e.shape[-1]), want.reshape(-1, want.shape[-1]))
        print(f'{case:18s} {prec}: peak-normalised err {O.peak_rel_err(e, want):.2e}   SI-SDR(ours, reference fp32) min {float(sis.min()):.1f} dB')
