"""Shard -> int16 H2D -> GPU widening -> ragged batches: bit-identical to feeding the float32 segments the reference's
datasets read (soundfile normalisation x / 32768)."""
import numpy as np
import pytest
import torch

import tss_with_dprnn_b200 as P
from tss_with_dprnn_b200 import shards
from tss_with_dprnn_b200.evaluate import evaluate

pytestmark = pytest.mark.gpu
KW = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
          n_repeats=1, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0)


def test_pcm16_widening_is_exact():
    pcm = torch.arange(-32768, 32768, dtype=torch.int32).to(torch.int16)
    out = torch.empty(pcm.numel(), device='cuda')
    P.lib().call('dprnn_pcm16_to_f32', pcm.cuda(), out, pcm.numel(), torch.cuda.current_stream().cuda_stream)
    assert torch.equal(out.cpu(), pcm.float() / 32768.0)


def test_evaluate_shard_matches_float_inputs(tmp_path):
    g = np.random.default_rng(3)
    items = []
    for i in range(7):
        T, Tr = int(g.integers(2000, 9000)), int(g.integers(1500, 6000))
        items.append({'id': f'u{i}', 'speaker': i, 'mixture': (3000 * g.standard_normal(T)).astype(np.int16),
                      'target': (2000 * g.standard_normal(T)).astype(np.int16),
                      'reference': (3000 * g.standard_normal(Tr)).astype(np.int16)})
    path = tmp_path / 'e.shard'
    shards.write_shard(path, items)
    torch.manual_seed(0)
    model = P.DPRNNSpeTasNet(**KW, fusion_type='cat').eval().cuda()
    got = shards.evaluate_shard(model, path, bucket=3, keep_audio=True)
    f32 = lambda k: [torch.from_numpy(it[k].astype(np.float32) / 32768.0) for it in items]
    want = evaluate(model, f32('mixture'), f32('reference'), f32('target'), bucket=3, keep_audio=True)
    assert [r['index'] for r in got] == [r['index'] for r in want] and len(got) == 7
    for a, b in zip(got, want):
        assert torch.equal(a['estimate'], b['estimate']) and a['si_sdr'] == b['si_sdr']
        assert a['id'] == items[a['index']]['id']
