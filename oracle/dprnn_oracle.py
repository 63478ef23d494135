"""CPU oracle for the DPRNN separation forward path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``tss_with_dprnn_b200``) never imports anything under ``oracle/`` and fails loudly when its CUDA
library is missing.

What it is: a functional (weights passed as a plain ``state_dict``) fp32 restatement, in
torch-on-CPU, of the algorithm the reference implements with ``nn.Module`` objects in
``src/models/{encoder_decoder,norms,dprnn,dprnn_spe,dprnn_spe_ira,dprnn_rawnet}.py``.  Integer work
(unfold / fold / nearest-upsample index maps) is restated in numpy.  Every function cites the
reference file:line it follows.  The reference itself ships no tests or golden vectors (SURVEY.md
section 4), so the oracle is pinned the other way the task allows: ``tests/golden/make_golden.py``
imports the live reference from ``/root/reference`` in the build container, runs it on seeded
inputs/weights and commits the outputs as fixtures; ``tests/test_oracle_vs_golden.py`` checks this
file against those fixtures (and, when ``/root/reference`` is importable, against the live modules).
Exception - parity unpinned: the RawNet3 sinc front-end (third-party ``asteroid_filterbanks==0.4.0``,
not vendored, not installed) is not restated here; cfg-4 style tests inject the ``[B,E]`` embedding.

Layout follows the reference (``[B, C, L]`` / ``[B, F, K, S]``) so per-stage tensors can be compared
with forward hooks on the reference modules.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class Config:
    """Constructor kwargs of the reference TasNet wrappers (dprnn.py:237-241, dprnn_spe.py:273-278)."""
    input_size: int = 64
    feature_size: int = 128
    hidden_size: int = 128
    chunk_length: int = 250
    kernel_size: int = 2
    hop_length: Optional[int] = 125
    n_repeats: int = 6
    bidirectional: bool = True
    norm_type: str = 'ln'
    activation_type: str = 'sigmoid'
    embeddings_size: int = 128
    num_spks: int = 251
    fusion_type: str = 'cat'
    stride: Optional[int] = None

    def __post_init__(self):
        if self.hop_length is None:
            self.hop_length = self.chunk_length // 2          # dprnn.py:127
        if self.stride is None:
            self.stride = self.kernel_size // 2               # dprnn.py:243


# --------------------------------------------------------------------------------------------
# integer index maps (bit-exact contracts)
# --------------------------------------------------------------------------------------------
def n_chunks(L: int, K: int, P: int) -> int:
    """Number of chunks F.unfold produces with kernel K, padding K both sides, stride P
    (dprnn.py:192-197): floor((L + 2K - K) / P) + 1."""
    return (L + K) // P + 1


def unfold_index_map(L: int, K: int, P: int) -> np.ndarray:
    """int64 [K, S] map: source frame t = s*P + k - K for chunk s, in-chunk position k; -1 where the
    source lies in the zero padding (dprnn.py:189-201)."""
    S = n_chunks(L, K, P)
    k = np.arange(K, dtype=np.int64)[:, None]
    s = np.arange(S, dtype=np.int64)[None, :]
    t = s * P + k - K
    t[(t < 0) | (t >= L)] = -1
    return t


def fold_coverage(L: int, K: int, P: int) -> np.ndarray:
    """How many (k, s) pairs land on each frame in F.fold (dprnn.py:203-217). 2 everywhere when P=K/2."""
    t = unfold_index_map(L, K, P)
    cov = np.zeros(L, dtype=np.int64)
    np.add.at(cov, t[t >= 0], 1)
    return cov


def nearest_upsample_index(L_in: int, L_out: int) -> np.ndarray:
    """Source index of nn.Upsample(size=L_out, mode='nearest') (dprnn_spe.py:181-182).  ATen computes
    scale = float(L_in) / L_out in fp32 and src = min(int(floorf(dst * scale)), L_in - 1), with
    shortcuts for equal sizes and exact 2x."""
    dst = np.arange(L_out, dtype=np.int64)
    if L_out == L_in:
        return dst
    if L_out == 2 * L_in:
        return dst >> 1
    scale = np.float32(L_in) / np.float32(L_out)
    src = np.floor(dst.astype(np.float32) * scale).astype(np.int64)
    return np.minimum(src, L_in - 1)


# --------------------------------------------------------------------------------------------
# stages
# --------------------------------------------------------------------------------------------
def encoder(x: Tensor, w: Tensor, stride: int = 1) -> Tensor:
    """Encoder.forward (encoder_decoder.py:25-33): Conv1d(1->N, k, stride, no bias) + ReLU. x [B,T]."""
    return F.relu(F.conv1d(x.unsqueeze(1), w, stride=stride))


def decoder(z: Tensor, w: Tensor, stride: int = 1) -> Tensor:
    """Decoder.forward (encoder_decoder.py:40-49): ConvTranspose1d(N->1) then squeeze to [B,T]."""
    y = F.conv_transpose1d(z, w, stride=stride)
    return y.reshape(z.shape[0], -1)


def chan_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Tensor:
    """nn.GroupNorm(1,C) (eps 1e-5) and norms.GlobLN (eps 1e-8, norms.py:6-31) are the same maths:
    per-sample mean / biased variance over every non-batch dim, per-channel affine."""
    dims = tuple(range(1, x.dim()))
    mean = x.mean(dim=dims, keepdim=True)
    var = x.var(dim=dims, keepdim=True, unbiased=False)
    shape = [1, -1] + [1] * (x.dim() - 2)
    return (x - mean) / torch.sqrt(var + eps) * gamma.view(shape) + beta.view(shape)


def norm_params(sd: SD, prefix: str, norm_type: str) -> Tuple[Tensor, Tensor, float]:
    """'gLN' modules name their parameters gamma/beta (norms.py:21-22), GroupNorm weight/bias."""
    if norm_type == 'gLN':
        return sd[prefix + '.gamma'], sd[prefix + '.beta'], 1e-8
    return sd[prefix + '.weight'], sd[prefix + '.bias'], 1e-5


def segmentation(x: Tensor, K: int, P: int) -> Tensor:
    """DPRNN._segmentation (dprnn.py:189-201) via the integer map. x [B,F,L] -> [B,F,K,S]."""
    B, Fd, L = x.shape
    idx = torch.from_numpy(unfold_index_map(L, K, P))
    xp = torch.cat([x, x.new_zeros(B, Fd, 1)], dim=-1)           # slot L holds the padding zero
    idx = torch.where(idx < 0, torch.full_like(idx, L), idx)
    return xp[:, :, idx]


def overlap_add(x: Tensor, L: int, K: int, P: int) -> Tensor:
    """DPRNN._overlap_add (dprnn.py:203-217). x [B2,F,K,S] -> [B2,F,L]; plain sum, no renormalisation."""
    B2, Fd, _, S = x.shape
    idx = torch.from_numpy(unfold_index_map(L, K, P)).reshape(-1)
    keep = idx >= 0
    out = x.new_zeros(B2, Fd, L)
    out.index_add_(2, idx[keep], x.reshape(B2, Fd, K * S)[:, :, keep])
    return out


def lstm_direction(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor,
                   reverse: bool) -> Tensor:
    """One direction of nn.LSTM (dprnn.py:23-28), explicit recurrence. Gate row order i,f,g,o;
    c' = sig(f) c + sig(i) tanh(g); h = sig(o) tanh(c'); zero initial state. x [N,T,I] -> [N,T,H]."""
    N, T, _ = x.shape
    H = w_hh.shape[1]
    gx = x @ w_ih.t() + (b_ih + b_hh)
    h = x.new_zeros(N, H)
    c = x.new_zeros(N, H)
    out = x.new_empty(N, T, H)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = gx[:, t] + h @ w_hh.t()
        i, f, gg, o = g.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


def lstm(x: Tensor, sd: SD, prefix: str, bidirectional: bool, fast: bool = True) -> Tensor:
    """SingleRNN.forward (dprnn.py:32-37). ``fast`` runs ATen's fused LSTM (what nn.LSTM calls) and is
    used for timing / large cases; ``fast=False`` is the explicit restatement. They are tested equal."""
    names = ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0']
    fw = [sd[f'{prefix}.{n}'] for n in names]
    bw = [sd[f'{prefix}.{n}_reverse'] for n in names] if bidirectional else []
    if fast:
        N = x.shape[0]
        H = fw[1].shape[1]
        D = 2 if bidirectional else 1
        z = x.new_zeros(D, N, H)
        out, _, _ = torch._VF.lstm(x, (z, z), fw + bw, True, 1, 0.0, False, bidirectional, True)
        return out
    outs = [lstm_direction(x, *fw, reverse=False)]
    if bidirectional:
        outs.append(lstm_direction(x, *bw, reverse=True))
    return torch.cat(outs, dim=-1)


def dprnn_block(x: Tensor, sd: SD, prefix: str, cfg: Config, fast: bool = True) -> Tensor:
    """DPRNNBlock.forward (dprnn.py:79-99). x [B,F,K,S]."""
    B, Fd, K, S = x.shape
    # intra-chunk: sequences along k, one per (b, s); always bidirectional (dprnn.py:58)
    seq = x.permute(0, 3, 2, 1).reshape(B * S, K, Fd)
    seq = lstm(seq, sd, f'{prefix}.intra_rnn.rnn', True, fast)
    seq = seq @ sd[f'{prefix}.intra_linear.weight'].t() + sd[f'{prefix}.intra_linear.bias']
    y = seq.reshape(B, S, K, Fd).permute(0, 3, 2, 1)
    g, b, eps = norm_params(sd, f'{prefix}.intra_norm', cfg.norm_type)
    x = x + chan_norm(y, g, b, eps)
    # inter-chunk: sequences along s, one per (b, k)
    seq = x.permute(0, 2, 3, 1).reshape(B * K, S, Fd)
    seq = lstm(seq, sd, f'{prefix}.inter_rnn.rnn', cfg.bidirectional, fast)
    seq = seq @ sd[f'{prefix}.inter_linear.weight'].t() + sd[f'{prefix}.inter_linear.bias']
    y = seq.reshape(B, K, S, Fd).permute(0, 3, 1, 2)
    g, b, eps = norm_params(sd, f'{prefix}.inter_norm', cfg.norm_type)
    return x + chan_norm(y, g, b, eps)


def conv1x1(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    """Pointwise Conv1d as a contraction over channels. x [B,Cin,L], w [Cout,Cin,1]."""
    y = torch.einsum('oc,bcl->bol', w[:, :, 0], x)
    return y if b is None else y + b.view(1, -1, 1)


def mask_head(x: Tensor, sd: SD, cfg: Config, L: int, pre: str = 'separation', fast: bool = True) -> Tensor:
    """Segmentation + blocks + PReLU + conv2d + overlap-add + gated head + activation
    (DPRNN.forward dprnn.py:166-187 == DPRNNSpe._dprnn_process dprnn_spe.py:231-248).
    x [B,F,L] (bottleneck output) -> masks [B,2,N,L]."""
    B = x.shape[0]
    Fd, K, P = cfg.feature_size, cfg.chunk_length, cfg.hop_length
    y = segmentation(x, K, P)
    for r in range(cfg.n_repeats):
        y = dprnn_block(y, sd, f'{pre}.dprnn_blocks.{r}', cfg, fast)
    a = sd[f'{pre}.prelu.weight']
    y = torch.where(y >= 0, y, a * y)
    S = y.shape[-1]
    y = torch.einsum('oc,bcks->boks', sd[f'{pre}.conv2d.weight'][:, :, 0, 0], y) \
        + sd[f'{pre}.conv2d.bias'].view(1, -1, 1, 1)
    y = y.reshape(B * 2, Fd, K, S)
    y = overlap_add(y, L, K, P)
    o = torch.tanh(conv1x1(y, sd[f'{pre}.out.0.weight'], sd[f'{pre}.out.0.bias']))
    g = torch.sigmoid(conv1x1(y, sd[f'{pre}.gate.0.weight'], sd[f'{pre}.gate.0.bias']))
    y = conv1x1(o * g, sd[f'{pre}.end_conv1x1.weight'])
    y = torch.sigmoid(y) if cfg.activation_type == 'sigmoid' else F.relu(y)
    return y.reshape(B, 2, cfg.input_size, L)


# ---- speaker branch ---------------------------------------------------------------------------
def batch_norm(y: Tensor, sd: SD, prefix: str, training: bool, new_stats: Optional[dict]) -> Tensor:
    """nn.BatchNorm1d (dprnn_spe.py:20-21): batch statistics (biased var) in train mode and a
    momentum-0.1 running update with the unbiased var; running statistics in eval mode. eps 1e-5."""
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    if training:
        mean = y.mean(dim=(0, 2))
        var = y.var(dim=(0, 2), unbiased=False)
        if new_stats is not None:
            n = y.shape[0] * y.shape[2]
            new_stats[prefix + '.running_mean'] = 0.9 * sd[prefix + '.running_mean'] + 0.1 * mean
            new_stats[prefix + '.running_var'] = 0.9 * sd[prefix + '.running_var'] + 0.1 * var * n / max(n - 1, 1)
            new_stats[prefix + '.num_batches_tracked'] = sd[prefix + '.num_batches_tracked'] + 1
    else:
        mean, var = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    return (y - mean.view(1, -1, 1)) / torch.sqrt(var.view(1, -1, 1) + 1e-5) * w.view(1, -1, 1) + b.view(1, -1, 1)


def res_block(x: Tensor, sd: SD, prefix: str, training: bool, new_stats: Optional[dict]) -> Tensor:
    """ResBlock.forward (dprnn_spe.py:31-42)."""
    y = conv1x1(x, sd[prefix + '.conv1.weight'])
    y = batch_norm(y, sd, prefix + '.batch_norm1', training, new_stats)
    y = torch.where(y >= 0, y, sd[prefix + '.prelu1.weight'] * y)
    y = conv1x1(y, sd[prefix + '.conv2.weight'])
    y = batch_norm(y, sd, prefix + '.batch_norm2', training, new_stats)
    skip = conv1x1(x, sd[prefix + '.conv_downsample.weight']) if (prefix + '.conv_downsample.weight') in sd else x
    y = y + skip
    y = torch.where(y >= 0, y, sd[prefix + '.prelu2.weight'] * y)
    Lp = y.shape[-1] // 3
    return y[..., :Lp * 3].reshape(y.shape[0], y.shape[1], Lp, 3).amax(dim=-1)     # MaxPool1d(3), floor


def speaker_embedding(aux: Tensor, aux_len: Tensor, sd: SD, cfg: Config, training: bool = False,
                      new_stats: Optional[dict] = None, pre: str = 'separation') -> Tensor:
    """DPRNNSpe._auxiliary (dprnn_spe.py:156-163): spk_encoder (dprnn_spe.py:115-122) then the time sum
    divided by a length derived from the *scalar* aux_len. aux [B,N,La] -> [B,E]."""
    p = f'{pre}.spk_encoder'
    y = chan_norm(aux, sd[p + '.0.weight'], sd[p + '.0.bias'], 1e-5)
    y = conv1x1(y, sd[p + '.1.weight'], sd[p + '.1.bias'])
    for i in (2, 3, 4):
        y = res_block(y, sd, f'{p}.{i}', training, new_stats)
    y = conv1x1(y, sd[p + '.5.weight'], sd[p + '.5.bias'])
    k = cfg.kernel_size
    aux_T = (aux_len - k) // (k // 2) + 1
    aux_T = ((aux_T // 3) // 3) // 3
    return y.sum(-1) / aux_T.reshape(-1, 1).float()


def fusion(e: Tensor, x: Tensor, sd: SD, cfg: Config, pre: str = 'separation') -> Tensor:
    """DPRNNSpe._fusion and helpers (dprnn_spe.py:165-229). e [B,E], x [B,N,L] (already normalised)."""
    ft = cfg.fusion_type
    L = x.shape[-1]

    def lin(name):
        return e @ sd[f'{pre}.{name}.weight'].t() + sd[f'{pre}.{name}.bias']

    if ft == 'cat':
        return torch.cat([x, e.unsqueeze(-1).expand(-1, -1, L)], dim=1)
    if ft == 'add':
        return x + lin('fusion_linear').unsqueeze(-1)
    if ft == 'mul':
        return x * lin('fusion_linear').unsqueeze(-1)
    if ft == 'film':
        return x * lin('fusion_linear_1').unsqueeze(-1) + lin('fusion_linear_2').unsqueeze(-1)
    if ft == 'att':
        k = cfg.kernel_size
        wa, ba = sd[f'{pre}.average.weight'], sd[f'{pre}.average.bias']
        La = (L - k) // k + 1
        frames = x[..., :La * k].reshape(x.shape[0], x.shape[1], La, k)
        avg = (frames * wa.view(1, -1, 1, k)).sum(-1) + ba.view(1, -1, 1)       # depthwise conv, stride k
        v = lin('fusion_linear')                                                 # [B,N]
        score = (avg * v.unsqueeze(-1)).sum(1)                                   # [B,La]
        sm = torch.softmax(score, dim=-1)
        att = sm.unsqueeze(1) * v.unsqueeze(-1) + v.unsqueeze(-1)                # [B,N,La]
        src = torch.from_numpy(nearest_upsample_index(La, L))
        return x * att[..., src]
    raise ValueError(ft)


# --------------------------------------------------------------------------------------------
# whole-model forwards
# --------------------------------------------------------------------------------------------
def tasnet_forward(mix: Tensor, sd: SD, cfg: Config, fast: bool = True) -> Tensor:
    """DPRNNTasNet.forward (dprnn.py:271-283): [B,T] -> [B,2,T]."""
    enc = encoder(mix, sd['encoder.conv1d.weight'], cfg.stride)
    g, b, eps = norm_params(sd, 'separation.bottleneck.0', cfg.norm_type)
    x = chan_norm(enc, g, b, eps)
    x = conv1x1(x, sd['separation.bottleneck.1.weight'], sd['separation.bottleneck.1.bias'])
    masks = mask_head(x, sd, cfg, enc.shape[-1], fast=fast)
    out = masks * enc.unsqueeze(1)
    return torch.stack([decoder(out[:, i], sd['decoder.weight'], cfg.stride) for i in range(2)], dim=1)


def spe_forward(mix: Tensor, ref: Tensor, ref_len: Tensor, sd: SD, cfg: Config, training: bool = False,
                new_stats: Optional[dict] = None, fast: bool = True,
                embedding: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """DPRNNSpeTasNet.forward (dprnn_spe.py:314-327) + DPRNNSpe.forward (dprnn_spe.py:125-154).
    ``embedding`` (cfg-4 style tests) replaces the speaker encoder output, as DPRNNRawNet does with
    RawNet3 (dprnn_rawnet.py:72-105)."""
    enc = encoder(mix, sd['encoder.conv1d.weight'], cfg.stride)
    if embedding is None:
        emb_in = encoder(ref, sd['encoder.conv1d.weight'], cfg.stride)
        e = speaker_embedding(emb_in, ref_len, sd, cfg, training, new_stats)
    else:
        e = embedding
    g, b, eps = norm_params(sd, 'separation.bottleneck.0', cfg.norm_type)
    x = chan_norm(enc, g, b, eps)
    x = fusion(e, x, sd, cfg)
    x = conv1x1(x, sd['separation.bottleneck.1.weight'], sd['separation.bottleneck.1.bias'])
    masks = mask_head(x, sd, cfg, enc.shape[-1], fast=fast)
    logits = e @ sd['separation.pred_linear.weight'].t() + sd['separation.pred_linear.bias']
    est = decoder((masks * enc.unsqueeze(1))[:, 0], sd['decoder.weight'], cfg.stride)
    return est, logits


def ira_forward(mix: Tensor, ref: Tensor, ref_len: Tensor, sd: SD, cfg: Config, training: bool = False,
                new_stats: Optional[dict] = None, fast: bool = True) -> Tuple[Tensor, Tensor]:
    """DPRNNSpeIRATasNet.forward (dprnn_spe_ira.py:179-190) + DPRNNSpeIRA.forward (:53-115): two masker
    passes over the same normalised encoding; the second embedding comes from the first estimate and is
    still divided by the *reference's* length (:84). In train mode the BatchNorm running statistics are
    updated twice (once per speaker-encoder call)."""
    enc = encoder(mix, sd['encoder.conv1d.weight'], cfg.stride)
    emb_in = encoder(ref, sd['encoder.conv1d.weight'], cfg.stride)
    L = enc.shape[-1]
    v0 = speaker_embedding(emb_in, ref_len, sd, cfg, training, new_stats)
    g, b, eps = norm_params(sd, 'separation.bottleneck.0', cfg.norm_type)
    xn = chan_norm(enc, g, b, eps)
    w1, b1 = sd['separation.bottleneck.1.weight'], sd['separation.bottleneck.1.bias']
    masks = mask_head(conv1x1(fusion(v0, xn, sd, cfg), w1, b1), sd, cfg, L, fast=fast)
    d0 = (masks * enc.unsqueeze(1))[:, 0]
    sd2 = sd if not new_stats else {**sd, **new_stats}
    v1 = speaker_embedding(d0, ref_len, sd2, cfg, training, new_stats)
    v1 = torch.cat([v0, v1], dim=1) @ sd['separation.aux_linear.weight'].t() + sd['separation.aux_linear.bias']
    masks = mask_head(conv1x1(fusion(v1, xn, sd, cfg), w1, b1), sd, cfg, L, fast=fast)
    d1 = (masks * enc.unsqueeze(1))[:, 0]
    logits = v1 @ sd['separation.pred_linear.weight'].t() + sd['separation.pred_linear.bias']
    return decoder(d1, sd['decoder.weight'], cfg.stride), logits


def peak_rel_err(a: Tensor, b: Tensor) -> float:
    """Error metric of SURVEY.md section 8c(iv): max|a-b| / max|b|."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def si_sdr_db(est: Tensor, target: Tensor) -> Tensor:
    """Scale-invariant SDR per row, zero-mean (the asteroid pairwise_neg_sisdr recipe the trainers use,
    trainer_spe.py:39; restated because asteroid is not installed - parity unpinned for this helper)."""
    est = est.double() - est.double().mean(-1, keepdim=True)
    target = target.double() - target.double().mean(-1, keepdim=True)
    s = (est * target).sum(-1, keepdim=True) * target / (target.pow(2).sum(-1, keepdim=True) + 1e-8)
    return 10 * torch.log10(s.pow(2).sum(-1) / ((est - s).pow(2).sum(-1) + 1e-8) + 1e-8)
