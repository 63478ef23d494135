"""CPU emulation of the tensor-core precision modes, to choose operand formats BEFORE spending GPU time.

Runs the oracle (oracle/dprnn_oracle.py) on a reference fixture with the roundings a tcgen05 mode applies:
LSTM operands x, h, [W_ih | W_hh] rounded to `--op` (bf16 | fp16 | tf32 | fp16x2 = W as hi+lo pair | fp32), fp32
accumulation, gate activations exact or with tanh.approx-sized noise (`--act approx`), the Linear after the LSTM with
`--lin` operands, the residual stream kept in fp32 (`--res fp32`) or in the operand format, the 1x1 convs with `--conv`
operands.  Prints the peak-normalised error against the reference fixture (north_star: <= 1e-3 in the fp32 mode).

    python tools/emulate_precision.py --case tasnet_r6_3s --op fp16 --act exact --lin fp16 --conv tf32
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_golden  # noqa: E402
from test_oracle_vs_golden import build_from_meta, oracle_cfg  # noqa: E402
from oracle import dprnn_oracle as O  # noqa: E402


def rnd(x, fmt):
    if fmt == 'fp32':
        return x
    if fmt == 'bf16':
        return x.to(torch.bfloat16).float()
    if fmt == 'fp16':
        return x.to(torch.float16).float()
    if fmt == 'tf32':       # 10 explicit mantissa bits, round to nearest even on the bit pattern
        i = x.contiguous().view(torch.int32)
        i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
        return i.view(torch.float32)
    raise ValueError(fmt)


def rnd_w(w, fmt):
    if fmt == 'fp16x2':     # hi + lo pair: two MMAs against the same activation operand
        hi = w.to(torch.float16).float()
        return hi + (w - hi).to(torch.float16).float()
    if fmt == 'bf16x2':
        hi = w.to(torch.bfloat16).float()
        return hi + (w - hi).to(torch.bfloat16).float()
    return rnd(w, fmt)


def act_fmt(fmt):
    return {'fp16x2': 'fp16', 'bf16x2': 'bf16'}.get(fmt, fmt)


class Emu:
    def __init__(self, a):
        self.a = a
        self.g = torch.Generator().manual_seed(7)

    def tanh(self, x):
        y = torch.tanh(x)
        if self.a.act == 'approx':      # tanh.approx.f32: max relative error 2^-10.987
            y = y * (1 + (torch.rand(y.shape, generator=self.g) * 2 - 1) * 2.0 ** -11)
        return y

    def sig(self, x):
        if self.a.act == 'approx':
            return 0.5 * self.tanh(0.5 * x) + 0.5
        return torch.sigmoid(x)

    def lstm_dir(self, xq, w_ih, w_hh, b, reverse):
        a = self.a
        N, T, _ = xq.shape
        H = w_hh.shape[1]
        wi, wh = rnd_w(w_ih, a.op), rnd_w(w_hh, a.op)
        h = xq.new_zeros(N, H)
        c = xq.new_zeros(N, H)
        out = xq.new_empty(N, T, H)
        gx = xq @ wi.t() + b
        for t in (range(T - 1, -1, -1) if reverse else range(T)):
            g = gx[:, t] + rnd(h, a.hop or act_fmt(a.op)) @ wh.t()
            i, f, gg, o = g.split(H, dim=1)
            c = self.sig(f) * c + self.sig(i) * self.tanh(gg)
            h = self.sig(o) * self.tanh(c)
            out[:, t] = h
        return out

    def lstm(self, x, sd, prefix, bid):
        xq = rnd(x, act_fmt(self.a.op))
        outs = []
        for sfx, rev in (('', False), ('_reverse', True))[:2 if bid else 1]:
            outs.append(self.lstm_dir(xq, sd[f'{prefix}.weight_ih_l0{sfx}'], sd[f'{prefix}.weight_hh_l0{sfx}'],
                                      sd[f'{prefix}.bias_ih_l0{sfx}'] + sd[f'{prefix}.bias_hh_l0{sfx}'], rev))
        return torch.cat(outs, -1)

    def linear(self, h, w, b):
        return rnd(h, act_fmt(self.a.lin)) @ rnd_w(w, self.a.lin).t() + b

    def block(self, x, sd, prefix, cfg):
        B, Fd, K, S = x.shape
        seq = x.permute(0, 3, 2, 1).reshape(B * S, K, Fd)
        seq = self.linear(self.lstm(seq, sd, f'{prefix}.intra_rnn.rnn', True), sd[f'{prefix}.intra_linear.weight'],
                          sd[f'{prefix}.intra_linear.bias'])
        if self.a.ybf:
            seq = rnd(seq, self.a.ybf)
        y = seq.reshape(B, S, K, Fd).permute(0, 3, 2, 1)
        g, b, eps = O.norm_params(sd, f'{prefix}.intra_norm', cfg.norm_type)
        x = x + O.chan_norm(y, g, b, eps)
        if self.a.res != 'fp32':
            x = rnd(x, self.a.res)
        seq = x.permute(0, 2, 3, 1).reshape(B * K, S, Fd)
        seq = self.linear(self.lstm(seq, sd, f'{prefix}.inter_rnn.rnn', cfg.bidirectional),
                          sd[f'{prefix}.inter_linear.weight'], sd[f'{prefix}.inter_linear.bias'])
        if self.a.ybf:
            seq = rnd(seq, self.a.ybf)
        y = seq.reshape(B, K, S, Fd).permute(0, 3, 1, 2)
        g, b, eps = O.norm_params(sd, f'{prefix}.inter_norm', cfg.norm_type)
        x = x + O.chan_norm(y, g, b, eps)
        if self.a.res != 'fp32':
            x = rnd(x, self.a.res)
        return x

    def conv(self, x, w, b=None):
        y = torch.einsum('oc,bcl->bol', rnd_w(w[:, :, 0], self.a.conv), rnd(x, act_fmt(self.a.conv)))
        return y if b is None else y + b.view(1, -1, 1)

    def mask_head(self, x, sd, cfg, L):
        B = x.shape[0]
        Fd, K, P = cfg.feature_size, cfg.chunk_length, cfg.hop_length
        y = O.segmentation(x, K, P)
        for r in range(cfg.n_repeats):
            y = self.block(y, sd, f'separation.dprnn_blocks.{r}', cfg)
        a = sd['separation.prelu.weight']
        y = torch.where(y >= 0, y, a * y)
        y2 = O.overlap_add(y, L, K, P)          # conv2d after the fold (A.6)
        w = sd['separation.conv2d.weight'][:, :, 0, 0]
        y = self.conv(y2, w[:, :, None]) + 2 * sd['separation.conv2d.bias'].view(1, -1, 1)
        y = y.reshape(B * 2, Fd, L)
        o = torch.tanh(self.conv(y, sd['separation.out.0.weight'], sd['separation.out.0.bias']))
        g = torch.sigmoid(self.conv(y, sd['separation.gate.0.weight'], sd['separation.gate.0.bias']))
        y = self.conv(o * g, sd['separation.end_conv1x1.weight'])
        y = torch.sigmoid(y) if cfg.activation_type == 'sigmoid' else torch.relu(y)
        return y.reshape(B, 2, cfg.input_size, L)

    def tasnet(self, mix, sd, cfg):
        enc = O.encoder(mix, sd['encoder.conv1d.weight'], cfg.stride)
        g, b, eps = O.norm_params(sd, 'separation.bottleneck.0', cfg.norm_type)
        x = O.chan_norm(enc, g, b, eps)
        x = self.conv(x, sd['separation.bottleneck.1.weight'], sd['separation.bottleneck.1.bias'])
        masks = self.mask_head(x, sd, cfg, enc.shape[-1])
        out = masks * enc.unsqueeze(1)
        return torch.stack([O.decoder(out[:, i], sd['decoder.weight'], cfg.stride) for i in range(2)], dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default='tasnet_r6_3s')
    ap.add_argument('--op', default='fp16')
    ap.add_argument('--act', default='exact', choices=['exact', 'approx'])
    ap.add_argument('--lin', default='fp16')
    ap.add_argument('--conv', default='tf32')
    ap.add_argument('--res', default='fp32')
    ap.add_argument('--ybf', default='')
    ap.add_argument('--hop', default='', help='format of the recurrent h operand (default: --op); fp32 = hi+lo split')
    ap.add_argument('--wscale', type=float, default=1.0, help='scale the LSTM weights (saturated-gate regime)')
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    meta, arr = load_golden(a.case)
    model = build_from_meta(meta).eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg = oracle_cfg(meta)
    mix = torch.from_numpy(arr['mix'])
    if a.wscale != 1.0:
        for k in sd:
            if '.rnn.weight_' in k:
                sd[k] = sd[k] * a.wscale
    with torch.no_grad():
        want = torch.from_numpy(arr['est']) if a.wscale == 1.0 else O.tasnet_forward(mix, sd, cfg)
        got = Emu(a).tasnet(mix, sd, cfg)
    sis = O.si_sdr_db(got.reshape(-1, got.shape[-1]), want.reshape(-1, want.shape[-1]))
    print(f'{a.case} op={a.op} act={a.act} lin={a.lin} conv={a.conv} res={a.res} ybf={a.ybf or "-"} wscale={a.wscale}: '
          f'peak-normalised err {O.peak_rel_err(got, want):.2e}  SI-SDR vs reference {float(sis.min()):.1f} dB', flush=True)


if __name__ == '__main__':
    main()
