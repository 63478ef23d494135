"""Packed evaluation / training shards: the on-disk format that feeds the ragged batcher.

The reference's datasets open one wav file per item and per field with ``soundfile.read(path, dtype='float32', start=,
stop=)`` (src/datasets/librimix_spe.py:50-55: mixture, source_1 = target, reference) and hand B = 1 items to the
inferencer.  Here a whole split is ONE file of 16-bit PCM that is memory-mapped:

    bytes 0..7    magic  b'DPRNNSH1'
    bytes 8..15   little-endian uint64: length J of the JSON header
    J bytes       JSON: {'sample_rate', 'fields': [...], 'utterances': [{'id', 'speaker', <field>: [offset, length], ...}]}
                  offsets / lengths in SAMPLES into the payload
    payload       int16 little-endian samples, 64-byte aligned start

``ShardReader`` hands out zero-copy int16 views; ``evaluate.evaluate`` packs a length bucket of them into one pinned
buffer, copies it to the device as int16 (half the bytes of float32) and widens it there with the exact soundfile
normalisation (x / 32768, ``dprnn_pcm16_to_f32``) - so the separated audio is bit-identical to feeding the float32
segments the reference reads.  ``build_shard_from_wavs`` converts the reference's csv-style item lists (16-bit PCM wav
paths + start / stop) with the standard library only.
"""
from __future__ import annotations

import json
import struct
import wave

import numpy as np
import torch

MAGIC = b'DPRNNSH1'
ALIGN = 64


def read_wav_pcm16(path, start: int = 0, stop: int | None = None) -> np.ndarray:
    """Mono 16-bit PCM wav -> int16 array of samples [start, stop) (what sf.read(..., start=, stop=) selects)."""
    with wave.open(str(path), 'rb') as w:
        if w.getsampwidth() != 2 or w.getnchannels() != 1:
            raise ValueError(f'{path}: expected mono 16-bit PCM (got {w.getnchannels()} ch, {8 * w.getsampwidth()} bit)')
        n = w.getnframes()
        stop = n if stop is None else min(int(stop), n)
        start = max(0, int(start))
        w.setpos(start)
        data = w.readframes(max(0, stop - start))
    return np.frombuffer(data, dtype='<i2').copy()


def write_shard(path, items, fields=('mixture', 'target', 'reference'), sample_rate: int = 8000):
    """items: iterable of dicts {'id': str, 'speaker': int, <field>: int16 array-like (1-D)}."""
    table, chunks, pos = [], [], 0
    for it in items:
        entry = {'id': str(it.get('id', len(table))), 'speaker': int(it.get('speaker', -1))}
        for f in fields:
            a = np.asarray(it[f])
            if a.dtype != np.int16 or a.ndim != 1:
                raise ValueError(f'field {f!r} must be a 1-D int16 array (16-bit PCM)')
            entry[f] = [pos, int(a.size)]
            chunks.append(a.astype('<i2', copy=False))
            pos += int(a.size)
        table.append(entry)
    header = json.dumps({'sample_rate': int(sample_rate), 'fields': list(fields), 'utterances': table}).encode()
    pre = len(MAGIC) + 8 + len(header)
    pad = (-pre) % ALIGN
    with open(path, 'wb') as f:
        f.write(MAGIC)
        f.write(struct.pack('<Q', len(header)))
        f.write(header)
        f.write(b'\0' * pad)
        for c in chunks:
            f.write(c.tobytes())
    return len(table)


def build_shard_from_wavs(path, rows, sample_rate: int = 8000):
    """rows: iterable of dicts with the reference's item description - 'mixture_path', 'source_1_path',
    'reference' (paths), 'start', 'stop', 'start_ref', 'stop_ref', 'speaker', 'id' (librimix_spe.py:41-62)."""
    def gen():
        for r in rows:
            yield {'id': r.get('id', ''), 'speaker': r.get('speaker', -1),
                   'mixture': read_wav_pcm16(r['mixture_path'], r.get('start', 0), r.get('stop')),
                   'target': read_wav_pcm16(r['source_1_path'], r.get('start', 0), r.get('stop')),
                   'reference': read_wav_pcm16(r['reference'], r.get('start_ref', 0), r.get('stop_ref'))}
    return write_shard(path, gen(), sample_rate=sample_rate)


class ShardReader:
    """Memory-mapped view of a shard: ``reader.field('mixture')`` is a list of zero-copy int16 torch tensors."""

    def __init__(self, path):
        with open(path, 'rb') as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise ValueError(f'{path}: not a DPRNN shard')
            (hlen,) = struct.unpack('<Q', f.read(8))
            self.meta = json.loads(f.read(hlen).decode())
        pre = len(MAGIC) + 8 + hlen
        self.payload_offset = pre + (-pre) % ALIGN
        self.data = np.memmap(path, dtype='<i2', mode='r', offset=self.payload_offset)
        self.utterances = self.meta['utterances']
        self.sample_rate = self.meta['sample_rate']
        need = max((u[f][0] + u[f][1] for u in self.utterances for f in self.meta['fields']), default=0)
        if need > self.data.size:
            raise ValueError(f'{path}: truncated payload ({self.data.size} samples, header needs {need})')

    def __len__(self):
        return len(self.utterances)

    def field(self, name):
        import warnings
        out = []
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')      # "array is not writable": the views are only ever read
            for u in self.utterances:
                off, n = u[name]
                out.append(torch.from_numpy(self.data[off:off + n]))
        return out

    def lengths(self, name='mixture'):
        return [u[name][1] for u in self.utterances]

    def speakers(self):
        return [u['speaker'] for u in self.utterances]

    def ids(self):
        return [u['id'] for u in self.utterances]


def evaluate_shard(model, path, bucket: int = 64, rank: int = 0, world: int = 1, keep_audio: bool = False):
    """The body of InferencerSpe.run / Inferencer.run over one shard: length-bucketed ragged batches, int16 H2D, SI-SDR on
    the GPU (evaluate.evaluate).  TSS models use 'mixture' / 'reference' / 'target'; DPRNNTasNet needs fields
    'mixture', 'source_1', 'source_2'."""
    from .evaluate import evaluate
    rd = ShardReader(path)
    if model.cfg['kind'] == 'bss':
        tg = [torch.cat([a, b]) for a, b in zip(rd.field('source_1'), rd.field('source_2'))]
        res = evaluate(model, rd.field('mixture'), None, tg, bucket=bucket, rank=rank, world=world, keep_audio=keep_audio)
    else:
        res = evaluate(model, rd.field('mixture'), rd.field('reference'), rd.field('target'), bucket=bucket, rank=rank,
                       world=world, keep_audio=keep_audio)
    ids = rd.ids()
    for r in res:
        r['id'] = ids[r['index']]
    return res
