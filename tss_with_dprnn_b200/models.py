"""Host-side mirror of the reference's model interface (``src/models``), backed by libdprnn_b200.

Same constructors / config keys, same ``state_dict`` layout (SURVEY.md Appendix A.1 - existing
``chkpts/*.pt`` strict-load), same ``forward`` signatures:

    DPRNNTasNet(mix[B,T])                         -> [B,2,T]            (src/models/dprnn.py:219-283)
    DPRNNSpeTasNet(mix, ref, ref_len)             -> (est[B,T], logits) (src/models/dprnn_spe.py:250-327)
    DPRNNSpeIRATasNet(mix, ref, ref_len)          -> (est, logits)      (src/models/dprnn_spe_ira.py:117-190)
    DPRNNRawNetTasNet(mix, ref16k)                -> (est, logits)      (src/models/dprnn_rawnet.py:107-182)

The ``nn`` sub-modules below are *parameter containers only* - they are created in the reference's
order so that ``torch.manual_seed(s)`` + default init reproduces the reference's weights bit for bit,
and so that ``load_state_dict`` sees identical keys.  No ``nn.Module.forward`` of theirs ever runs:
``forward`` hands raw device pointers to the CUDA library through :mod:`.engine`.  There is no CPU
path and no PyTorch fallback.
"""
from __future__ import annotations

import torch
from torch import nn

from .engine import Engine
from .rawnet import RawNet3


class _GlobLN(nn.Module):
    """Parameter container for norms.GlobLN (src/models/norms.py:17-31): keys ``gamma`` / ``beta``."""

    def __init__(self, channels: int):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))


def _norm(norm_type: str, channels: int, strict: bool) -> nn.Module:
    # DPRNN / DPRNNBlock accept exactly 'gLN' | 'ln' (dprnn.py:72-77,130-133); DPRNNSpe treats anything
    # that is not 'gLN' as GroupNorm (dprnn_spe.py:108-111).
    if norm_type == 'gLN':
        return _GlobLN(channels)
    if norm_type == 'ln' or not strict:
        return nn.GroupNorm(1, channels)
    raise ValueError(f"norm_type must be 'gLN' or 'ln', got {norm_type!r}")


class _RNN(nn.Module):
    """Container matching SingleRNN (dprnn.py:7-37): one nn.LSTM under the attribute ``rnn``."""

    def __init__(self, input_size, hidden_size, bidirectional):
        super().__init__()
        self.rnn = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=1, batch_first=True,
                           bidirectional=bidirectional)


class _Block(nn.Module):
    """Container matching DPRNNBlock (dprnn.py:39-77)."""

    def __init__(self, feature_size, hidden_size, norm_type, bidirectional):
        super().__init__()
        self.intra_rnn = _RNN(feature_size, hidden_size, True)
        self.intra_linear = nn.Linear(2 * hidden_size, feature_size)
        self.inter_rnn = _RNN(feature_size, hidden_size, bidirectional)
        self.inter_linear = nn.Linear((2 if bidirectional else 1) * hidden_size, feature_size)
        self.intra_norm = _norm(norm_type, feature_size, True)
        self.inter_norm = _norm(norm_type, feature_size, True)


class _ResBlock(nn.Module):
    """Container matching ResBlock (dprnn_spe.py:8-29)."""

    def __init__(self, in_dims, out_dims):
        super().__init__()
        self.conv1 = nn.Conv1d(in_dims, out_dims, 1, bias=False)
        self.conv2 = nn.Conv1d(out_dims, out_dims, 1, bias=False)
        self.batch_norm1 = nn.BatchNorm1d(out_dims)
        self.batch_norm2 = nn.BatchNorm1d(out_dims)
        self.prelu1 = nn.PReLU()
        self.prelu2 = nn.PReLU()
        if in_dims != out_dims:
            self.conv_downsample = nn.Conv1d(in_dims, out_dims, 1, bias=False)


class _Masker(nn.Module):
    """Container for DPRNN / DPRNNSpe / DPRNNSpeIRA / DPRNNRawNet ("separation.*" keys)."""

    def __init__(self, kind, input_size, feature_size, hidden_size, chunk_length, hop_length, n_repeats,
                 bidirectional, norm_type, activation_type, O=128, P=256, embeddings_size=128, num_spks=251,
                 kernel_size=2, fusion_type='cat'):
        super().__init__()
        N, F = input_size, feature_size
        # --- DPRNN.__init__ (dprnn.py:116-164) ---
        first_norm = _norm(norm_type, N, True)
        self.bottleneck = nn.Sequential(first_norm, nn.Conv1d(N, F, 1))
        self.dprnn_blocks = nn.Sequential(*[_Block(F, hidden_size, norm_type, bidirectional)
                                            for _ in range(n_repeats)])
        self.prelu = nn.PReLU()
        self.conv2d = nn.Conv2d(F, 2 * F, kernel_size=1)
        self.out = nn.Sequential(nn.Conv1d(F, F, 1), nn.Tanh())
        self.gate = nn.Sequential(nn.Conv1d(F, F, 1), nn.Sigmoid())
        self.end_conv1x1 = nn.Conv1d(F, N, 1, bias=False)
        if activation_type not in ('sigmoid', 'relu'):
            raise ValueError(f"activation_type must be 'sigmoid' or 'relu', got {activation_type!r}")
        if kind == 'bss':
            return
        # --- DPRNNSpe.__init__ (dprnn_spe.py:85-123) ---
        E = embeddings_size
        if fusion_type not in ('cat', 'add', 'mul', 'film', 'att'):
            raise ValueError(f'unknown fusion_type {fusion_type!r}')
        if fusion_type == 'cat':
            start = nn.Conv1d(N + E, F, 1)
        elif fusion_type in ('add', 'mul'):
            self.fusion_linear = nn.Linear(E, N)
            start = nn.Conv1d(N, F, 1)
        elif fusion_type == 'film':
            self.fusion_linear_1 = nn.Linear(E, N)
            self.fusion_linear_2 = nn.Linear(E, N)
            start = nn.Conv1d(N, F, 1)
        else:  # att: frozen depthwise averaging conv (dprnn_spe.py:99-104)
            self.fusion_linear = nn.Linear(E, N)
            self.average = nn.Conv1d(N, N, kernel_size, kernel_size, groups=N)
            self.average.weight = nn.Parameter(torch.ones(N, 1, kernel_size) / kernel_size, requires_grad=False)
            self.average.bias = nn.Parameter(torch.zeros(N), requires_grad=False)
            start = nn.Conv1d(N, F, 1)
        self.bottleneck = nn.Sequential(_norm(norm_type, N, False), start)
        self.spk_encoder = nn.Sequential(
            nn.GroupNorm(1, N), nn.Conv1d(N, O, 1), _ResBlock(O, O), _ResBlock(O, P), _ResBlock(P, P),
            nn.Conv1d(P, E, 1))
        self.pred_linear = nn.Linear(E, num_spks)
        if kind == 'ira':
            self.aux_linear = nn.Linear(2 * E, E)      # dprnn_spe_ira.py:51
        if kind == 'rawnet':
            # DPRNNRawNet.__init__ builds the whole DPRNNSpe first (the ResNet above consumes the RNG exactly as in the
            # reference) and then replaces the speaker encoder with RawNet3 (dprnn_rawnet.py:57-70)
            self.spk_encoder = RawNet3(nOut=E)


class _TasNetBase(nn.Module):
    kind = 'bss'

    def __init__(self, input_size, feature_size=128, hidden_size=128, chunk_length=200, kernel_size=2,
                 hop_length=None, n_repeats=6, bidirectional=True, rnn_type='LSTM', norm_type='ln',
                 activation_type='sigmoid', dropout=0, stride=None, **spe):
        super().__init__()
        if rnn_type != 'LSTM':
            raise NotImplementedError("only rnn_type='LSTM' is built (every shipped config uses it)")
        if dropout != 0:
            raise NotImplementedError('dropout must be 0 (a single-layer nn.LSTM ignores it anyway)')
        self.stride = stride if stride is not None else kernel_size // 2
        self.cfg = dict(input_size=input_size, feature_size=feature_size, hidden_size=hidden_size,
                        chunk_length=chunk_length, kernel_size=kernel_size,
                        hop_length=hop_length if hop_length is not None else chunk_length // 2,
                        n_repeats=n_repeats, bidirectional=bidirectional, norm_type=norm_type,
                        activation_type=activation_type, stride=self.stride, kind=self.kind,
                        embeddings_size=spe.get('embeddings_size', 128), num_spks=spe.get('num_spks', 251),
                        fusion_type=spe.get('fusion_type', 'cat') if self.kind != 'bss' else None)
        # Encoder (encoder_decoder.py:14-23): key 'encoder.conv1d.weight'
        self.encoder = nn.Module()
        self.encoder.conv1d = nn.Conv1d(1, input_size, kernel_size, stride=self.stride, bias=False)
        self.separation = _Masker(self.kind, input_size, feature_size, hidden_size, chunk_length,
                                  self.cfg['hop_length'], n_repeats, bidirectional, norm_type, activation_type,
                                  kernel_size=kernel_size, **spe)
        # Decoder (encoder_decoder.py:35-38): key 'decoder.weight'
        self.decoder = nn.ConvTranspose1d(input_size, 1, kernel_size, stride=self.stride, bias=False)
        self._engine = Engine(self)

    #: 'fp32' = exact fp32 everywhere; 'bf16' / 'fp16' = tcgen05 LSTM / Linear contractions with bf16 / fp16 operands
    #: (fp32 accumulation); 'fp16' stays within 1e-3 (peak-normalised) of the reference's fp32 path at the same speed
    @property
    def precision(self) -> str:
        return self._engine.precision

    @precision.setter
    def precision(self, mode: str):
        self._engine.set_precision(mode)

    def invalidate(self):
        """Forget the kernel-layout weight copies and the captured CUDA graphs.  Needed only after writing parameters or
        buffers in a way torch cannot see: through ``p.data`` (EMA, weight clipping, ``p.data.copy_``) or raw pointers -
        neither bumps the tensors' version counters the caches are keyed on.  ``load_state_dict``, ``.to()`` / ``.cuda()``
        / ``.half()`` (``_apply``) and the fused optimiser step of ``SpeTrainStep`` call it themselves."""
        self._engine.invalidate()

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        eng = self.__dict__.get('_engine')
        if eng is not None:
            eng.invalidate()
        return out

    def load_state_dict(self, *a, **kw):
        out = super().load_state_dict(*a, **kw)
        self._engine.invalidate()
        return out

    #: number of concurrent CUDA streams the batch is split over inside forward (results do not depend on it)
    @property
    def n_streams(self) -> int:
        return self._engine.n_streams

    @n_streams.setter
    def n_streams(self, n: int):
        self._engine.n_streams = max(1, int(n))


class DPRNNTasNet(_TasNetBase):
    """Blind separation of two speakers; mirrors src/models/dprnn.py:219-283."""
    kind = 'bss'

    def forward(self, input):
        return self._engine.forward_bss(input)

    def forward_ragged(self, inputs):
        """Batch of utterances of DIFFERENT lengths: list of [T_b] -> list of [2, T_b]; every result equals
        ``forward(inputs[b][None])[0]`` (the reference's B = 1 test loop, src/inferencers/inferencer.py:54-71)."""
        return self._engine.forward_bss_ragged(inputs)


class DPRNNSpeTasNet(_TasNetBase):
    """Target-speaker separation with a ResNet speaker encoder; mirrors src/models/dprnn_spe.py:250-327."""
    kind = 'spe'

    def __init__(self, input_size, feature_size=128, hidden_size=128, chunk_length=200, kernel_size=2,
                 hop_length=None, n_repeats=6, bidirectional=True, rnn_type='LSTM', norm_type='gLN',
                 activation_type='sigmoid', dropout=0, stride=None, O=128, P=256, embeddings_size=128,
                 num_spks=251, fusion_type='cat'):
        super().__init__(input_size, feature_size, hidden_size, chunk_length, kernel_size, hop_length, n_repeats,
                         bidirectional, rnn_type, norm_type, activation_type, dropout, stride, O=O, P=P,
                         embeddings_size=embeddings_size, num_spks=num_spks, fusion_type=fusion_type)

    def forward(self, input, aux, aux_len):
        return self._engine.forward_spe(input, aux, aux_len)

    def forward_ragged(self, inputs, auxs):
        """Batch of utterances of DIFFERENT lengths: lists of [T_b] mixtures and [Tr_b] references ->
        (list of [T_b] estimates, logits [B, num_spks]); utterance b gets ``forward(inputs[b][None], auxs[b][None],
        tensor(float(Tr_b)))`` of eval() mode - the reference's B = 1 test loop (src/inferencers/inferencer_spe.py:25-45)."""
        return self._engine.forward_spe_ragged(inputs, auxs)

    def forward_with_embedding(self, input, embedding):
        """Masker + decoder with an externally supplied speaker embedding [B,E] (what DPRNNRawNet does
        with RawNet3's output, dprnn_rawnet.py:72-105)."""
        return self._engine.forward_spe(input, None, None, embedding=embedding)


class DPRNNSpeIRATasNet(DPRNNSpeTasNet):
    """Two masker passes with re-embedding of the first estimate; mirrors src/models/dprnn_spe_ira.py:117-190."""
    kind = 'ira'

    def forward(self, input, aux, aux_len):
        return self._engine.forward_ira(input, aux, aux_len)

    def forward_ragged(self, inputs, auxs):
        return self._engine.forward_ira_ragged(inputs, auxs)


class DPRNNRawNetTasNet(_TasNetBase):
    """Target-speaker separation with a RawNet3 speaker encoder on the raw 16 kHz reference; mirrors
    src/models/dprnn_rawnet.py:107-182.  ``forward(input, aux)`` has no ``aux_len`` (dprnn_rawnet.py:171).
    The masker, fusion and decoder are the CUDA path; RawNet3 itself is the stage-1 GPU restatement of rawnet.py."""
    kind = 'rawnet'

    def __init__(self, input_size, feature_size=128, hidden_size=128, chunk_length=200, kernel_size=2,
                 hop_length=None, n_repeats=6, bidirectional=True, rnn_type='LSTM', norm_type='gLN',
                 activation_type='sigmoid', dropout=0, stride=None, O=128, P=256, embeddings_size=128,
                 num_spks=251, fusion_type='cat'):
        super().__init__(input_size, feature_size, hidden_size, chunk_length, kernel_size, hop_length, n_repeats,
                         bidirectional, rnn_type, norm_type, activation_type, dropout, stride, O=O, P=P,
                         embeddings_size=embeddings_size, num_spks=num_spks, fusion_type=fusion_type)

    def forward(self, input, aux):
        se = self.separation.spk_encoder
        se.allow_tf32 = self.precision != 'fp32'
        if self.training and torch.is_grad_enabled():
            # TrainerRawNet (src/trainers/trainer_rawnet.py:31-56): RawNet3 as library ops under torch autograd, the masker
            # + decoder as the hand-written autograd node that also returns d loss / d embedding (train.EmbTrainFunction)
            emb = se.embed_autograd(aux)
        else:
            emb = se.embed(aux)
        return self._engine.forward_spe(input, None, None, embedding=emb)

    def forward_with_embedding(self, input, embedding):
        return self._engine.forward_spe(input, None, None, embedding=embedding)
