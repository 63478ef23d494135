"""Batched evaluation driver: what replaces the B = 1 Python loops of
``InferencerSpe.run`` / ``Inferencer.run`` (src/inferencers/inferencer_spe.py:25-45, inferencer.py:54-71).

Utterances are sorted into length buckets, every bucket runs as ONE ragged batch through ``model.forward_ragged``
(each result is bit-identical to the reference-style per-utterance call), SI-SDR is computed on the GPU and only the
separated audio (optional) and one float per utterance come back to the host.  Host -> device copies of bucket i+1 are
issued from pinned memory on a side stream while bucket i computes.  PESQ / STOI stay with the caller (CPU libraries).
"""
from __future__ import annotations

import torch

from ._lib import lib
from .sharding import length_buckets, lpt_assign, chunk_count


class _Pcm:
    """A pinned int16 host buffer whose .to(device) lands as float32 (x / 32768) through dprnn_pcm16_to_f32."""

    def __init__(self, host):
        self.host = host

    def to(self, dev, non_blocking=True):
        raw = self.host.to(dev, non_blocking=non_blocking)
        out = torch.empty(raw.numel(), device=dev, dtype=torch.float32)
        lib().call('dprnn_pcm16_to_f32', raw, out, raw.numel(), torch.cuda.current_stream().cuda_stream)
        raw.record_stream(torch.cuda.current_stream())
        return out


def si_sdr(est: torch.Tensor, target: torch.Tensor, lengths=None) -> torch.Tensor:
    """SI-SDR in dB per utterance on the GPU.  est / target: [B, T], or packed 1-D tensors with `lengths`."""
    if not est.is_cuda:
        raise RuntimeError('si_sdr: CUDA tensors required (no CPU path)')
    est, target = est.contiguous().float(), target.contiguous().float()
    st = torch.cuda.current_stream().cuda_stream
    if lengths is None:
        B, T = est.shape
        out = torch.empty(B, device=est.device)
        lib().call('dprnn_si_sdr', est, target, None, None, T, B, out, st)
        return out
    lens = torch.tensor([int(n) for n in lengths], dtype=torch.int64)
    off = torch.cumsum(lens, 0) - lens
    out = torch.empty(len(lengths), device=est.device)
    lib().call('dprnn_si_sdr', est, target, off.to(est.device), lens.to(est.device), 0, len(lengths), out, st)
    return out


def evaluate(model, mixtures, references=None, targets=None, bucket: int = 64, rank: int = 0, world: int = 1,
             keep_audio: bool = False):
    """Separate a list of utterances (1-D CPU float tensors of any lengths) with `model` (a DPRNNTasNet /
    DPRNNSpeTasNet / DPRNNSpeIRATasNet on a CUDA device, eval mode).

    Returns a list of dicts ``{'index', 'si_sdr' (if targets), 'logits' (TSS models), 'estimate' (if keep_audio)}`` for
    the utterances of this rank: buckets are assigned to the `world` ranks with the LPT rule, no communication."""
    dev = next(model.parameters()).device
    tss = references is not None
    lengths = [int(m.numel()) for m in mixtures]
    buckets = length_buckets(lengths, bucket)
    mine = lpt_assign([sum(chunk_count(lengths[i]) for i in b) for b in buckets], world)[rank]
    copy_stream = torch.cuda.Stream(device=dev)

    def stage(idx):
        """pinned packed host buffers -> device, on the copy stream"""
        def pack(items):
            if items[idx[0]].dtype == torch.int16:        # 16-bit PCM (shards.py): half the H2D bytes, widened on the GPU
                return _Pcm(torch.cat([items[i].reshape(-1) for i in idx]).pin_memory())
            flat = torch.cat([items[i].reshape(-1).float() for i in idx]).pin_memory()
            return flat
        with torch.cuda.stream(copy_stream):
            out = dict(idx=idx, mix=pack(mixtures).to(dev, non_blocking=True), Ts=[lengths[i] for i in idx])
            if tss:
                out['ref'] = pack(references).to(dev, non_blocking=True)
                out['Trs'] = [int(references[i].numel()) for i in idx]
            if targets is not None:
                out['tgt'] = pack(targets).to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            out['ready'] = ev
        return out

    results = []
    nxt = stage(buckets[mine[0]]) if mine else None
    with torch.no_grad():
        for j in range(len(mine)):
            cur = nxt
            nxt = stage(buckets[mine[j + 1]]) if j + 1 < len(mine) else None
            main = torch.cuda.current_stream()
            main.wait_event(cur['ready'])
            for v in cur.values():          # allocated on the copy stream, consumed on this one: keep the caching allocator
                t = v.data if isinstance(v, _Pcm) else v        # from re-using the blocks before this stream is done
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(main)
            if tss:
                est, logits = model.forward_ragged((cur['mix'], cur['Ts']), (cur['ref'], cur['Trs']))
            else:
                est, logits = model.forward_ragged((cur['mix'], cur['Ts'])), None
            scores = None
            if targets is not None:
                if tss:
                    scores = si_sdr(torch.cat(est), cur['tgt'], cur['Ts']).cpu()
                else:       # BSS: best permutation of the two estimates per utterance (PIT, inferencer.py:60)
                    e0, e1 = torch.cat([e[0] for e in est]), torch.cat([e[1] for e in est])
                    tg = cur['tgt'].view(-1)
                    t0 = torch.cat([t[:n] for t, n in zip(tg.split([2 * n for n in cur['Ts']]), cur['Ts'])])
                    t1 = torch.cat([t[n:] for t, n in zip(tg.split([2 * n for n in cur['Ts']]), cur['Ts'])])
                    a = (si_sdr(e0, t0, cur['Ts']) + si_sdr(e1, t1, cur['Ts'])) / 2
                    b = (si_sdr(e0, t1, cur['Ts']) + si_sdr(e1, t0, cur['Ts'])) / 2
                    scores = torch.maximum(a, b).cpu()
            for k, i in enumerate(cur['idx']):
                r = {'index': i}
                if scores is not None:
                    r['si_sdr'] = float(scores[k])
                if logits is not None:
                    r['logits'] = logits[k].cpu()
                if keep_audio:
                    r['estimate'] = est[k].cpu()
                results.append(r)
    return results
