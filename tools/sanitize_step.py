"""Small forwards through every code path (uniform + ragged, fp32 + bf16, att + cat fusion, IRA, TasNet) for
`compute-sanitizer --tool memcheck python tools/sanitize_step.py` (one tool per gpurun call, small shapes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tss_with_dprnn_b200 as P  # noqa: E402

KW = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
          n_repeats=1, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0)
torch.manual_seed(0)
g = torch.Generator().manual_seed(1)
w = lambda n: (0.05 * torch.randn(n, generator=g)).cuda()
with torch.no_grad():
    for prec in ('bf16', 'fp32'):
        for cls, kw in ((P.DPRNNSpeTasNet, dict(fusion_type='att')), (P.DPRNNSpeTasNet, dict(fusion_type='cat')),
                        (P.DPRNNSpeIRATasNet, dict(fusion_type='film')), (P.DPRNNTasNet, {})):
            m = cls(**KW, **kw).eval().cuda()
            m.precision = prec
            m._engine.use_graphs = False
            mix, ref = torch.stack([w(3000), w(3000)]), torch.stack([w(2000), w(2000)])
            if cls is P.DPRNNTasNet:
                out = m(mix)
                rag = m.forward_ragged([w(1500), w(2777)])
            else:
                out = m(mix, ref, torch.tensor(2000.))[0]
                rag = m.forward_ragged([w(1500), w(2777)], [w(900), w(1300)])[0]
            torch.cuda.synchronize()
            print(prec, cls.__name__, kw, 'ok', float(out.abs().mean()))
print('done')
