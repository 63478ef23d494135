"""The packed shard format (tss_with_dprnn_b200/shards.py): write / memory-mapped read round trip, wav import with
start/stop crops as the reference's datasets select them, and the bucket -> rank assignment covering every utterance once."""
import wave

import numpy as np
import pytest
import torch

from tss_with_dprnn_b200 import shards
from tss_with_dprnn_b200.sharding import length_buckets, lpt_assign, chunk_count


def _items(n, seed=0):
    g = np.random.default_rng(seed)
    out = []
    for i in range(n):
        T, Tr = int(g.integers(300, 4000)), int(g.integers(300, 2000))
        out.append({'id': f'utt{i}', 'speaker': int(g.integers(0, 251)),
                    'mixture': g.integers(-32768, 32767, T, dtype=np.int16),
                    'target': g.integers(-32768, 32767, T, dtype=np.int16),
                    'reference': g.integers(-32768, 32767, Tr, dtype=np.int16)})
    return out


def test_shard_round_trip(tmp_path):
    items = _items(17)
    path = tmp_path / 'test.shard'
    assert shards.write_shard(path, items) == 17
    rd = shards.ShardReader(path)
    assert len(rd) == 17 and rd.sample_rate == 8000 and rd.payload_offset % 64 == 0
    for f in ('mixture', 'target', 'reference'):
        got = rd.field(f)
        for it, t in zip(items, got):
            assert t.dtype == torch.int16 and np.array_equal(t.numpy(), it[f])
    assert rd.lengths() == [it['mixture'].size for it in items]
    assert rd.speakers() == [it['speaker'] for it in items] and rd.ids() == [it['id'] for it in items]
    with open(path, 'r+b') as f:                         # a truncated file is detected
        f.truncate(rd.payload_offset + 100)
    with pytest.raises(ValueError):
        shards.ShardReader(path)


def test_wav_import_with_crops(tmp_path):
    g = np.random.default_rng(1)
    pcm = {n: g.integers(-20000, 20000, 5000, dtype=np.int16) for n in ('mix', 's1', 'ref')}
    for n, a in pcm.items():
        with wave.open(str(tmp_path / f'{n}.wav'), 'wb') as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(8000)
            w.writeframes(a.tobytes())
    rows = [{'id': 'a_b', 'speaker': 3, 'mixture_path': tmp_path / 'mix.wav', 'source_1_path': tmp_path / 's1.wav',
             'reference': tmp_path / 'ref.wav', 'start': 100, 'stop': 3100, 'start_ref': 7, 'stop_ref': 2007}]
    shards.build_shard_from_wavs(tmp_path / 'w.shard', rows)
    rd = shards.ShardReader(tmp_path / 'w.shard')
    assert np.array_equal(rd.field('mixture')[0].numpy(), pcm['mix'][100:3100])
    assert np.array_equal(rd.field('target')[0].numpy(), pcm['s1'][100:3100])
    assert np.array_equal(rd.field('reference')[0].numpy(), pcm['ref'][7:2007])
    # the float32 the reference would see: soundfile normalises 16-bit PCM by 1 / 32768 (exact in float32)
    f32 = rd.field('mixture')[0].float() / 32768.0
    assert float(f32.abs().max()) <= 1.0 and f32.dtype == torch.float32


def test_bucket_assignment_covers_the_shard(tmp_path):
    items = _items(50, seed=2)
    shards.write_shard(tmp_path / 'b.shard', items)
    rd = shards.ShardReader(tmp_path / 'b.shard')
    lengths = rd.lengths()
    buckets = length_buckets(lengths, 8)
    seen = []
    for rank in range(3):
        mine = lpt_assign([sum(chunk_count(lengths[i]) for i in b) for b in buckets], 3)[rank]
        seen += [i for j in mine for i in buckets[j]]
    assert sorted(seen) == list(range(50))
