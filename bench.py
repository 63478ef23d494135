#!/usr/bin/env python
"""Headline benchmark: separated audio-seconds per second, DPRNN-Spe (cat fusion), 3 s @ 8 kHz.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on this box's cores

A "step" is one forward pass of the separation path over one batch of synthetic mixtures
(BASELINE.json configs[1]: DPRNN-Spe cat, 3-s mix + 3-s reference at 8 kHz, batch 64 per GPU, weak scaling:
utterances are independent, so each rank runs its own batch and there is no data-path collective).
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR = 8000
KW = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
          n_repeats=6, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0, fusion_type='cat')
METRIC = 'separated audio-sec/sec (DPRNN-Spe, 3 s @8 kHz)'

# Algorithmic work of one 3-s utterance: the LSTM recurrence h_{t-1} W_hh^T of all
# 12 RNN layers (2 directions): 2 * 128 * 512 flop per position and direction.
POS_PER_UTT = 250 * 194
RECUR_FLOP_PER_UTT = 12 * 2 * POS_PER_UTT * 2 * 128 * 512           # 152.6 GFLOP
PROJ_FLOP_PER_UTT = RECUR_FLOP_PER_UTT                               # x W_ih^T, same shape


def synth(batch, T, rank, device='cpu'):
    g = torch.Generator().manual_seed(1234 + 1000 * rank)
    mix = 0.05 * torch.randn(batch, T, generator=g)
    g = torch.Generator().manual_seed(1235 + 1000 * rank)
    ref = 0.05 * torch.randn(batch, T, generator=g)
    return mix.to(device), ref.to(device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'power_w_max': max(pw),
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own classes (baseline/_ref, installed unmodified) when present,
# else the oracle port (torch CPU restatement calling the same ATen LSTM the reference calls)
# ----------------------------------------------------------------------------------------------------------
def cpu_forward_fn():
    ref_root = os.path.join(ROOT, 'baseline', '_ref')
    torch.manual_seed(0)
    if os.path.isdir(os.path.join(ref_root, 'src', 'models')):
        sys.path.insert(0, ref_root)
        from src.models.dprnn_spe import DPRNNSpeTasNet as RefModel
        model = RefModel(**KW).eval()
        return 'reference', (lambda mix, ref, rl: model(mix, ref, rl)[0])
    from oracle import dprnn_oracle as O
    import tss_with_dprnn_b200 as P
    sd = P.DPRNNSpeTasNet(**KW).state_dict()
    cfg = O.Config(fusion_type='cat')
    return 'port', (lambda mix, ref, rl: O.spe_forward(mix, ref, rl, sd, cfg)[0])


def time_cpu(batch, T, steps, warmup):
    kind, fwd = cpu_forward_fn()
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    mix, ref = synth(batch, T, 0)
    rl = torch.tensor(float(T))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fwd(mix, ref, rl)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return kind, cores, times


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    batch = args.cpu_batch
    kind, cores, times = time_cpu(batch, args.samples, args.steps, args.warmup)
    total = sum(times)
    value = batch * args.samples / SR * len(times) / total
    sample = f'{batch} x {args.samples / SR:g}-s utterance(s) per step (bounded sample of the batch-{args.batch} workload)'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'audio-s/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args),
        'cpu_baseline': {'value': value, 'unit': 'audio-s/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {'workload': f'cfg2: DPRNN-Spe (cat) TSS inference, {args.samples / SR:g}-s mix + reference @ 8 kHz, '
                        f'batch {args.batch} per GPU, full depth (6 blocks)',
            'batch_per_gpu': args.batch, 'samples': args.samples, 'precision': args.precision,
            'streams': args.streams,
            'l2': 'no flush needed: each step streams >10 GB of intermediates through a 126 MB L2',
            'parallelism': f'utterance sharding x{args.gpus}, no collective'}


# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import tss_with_dprnn_b200 as P
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(0)
    model = P.DPRNNSpeTasNet(**KW).eval().to(dev)      # same seeded weights on every rank (replicated, 16 MB)
    model.precision = args.precision
    model.n_streams = args.streams
    if args.fused_tail is not None:
        model._engine.fused_tail = bool(args.fused_tail)
    B, T = args.batch, args.samples
    mix_h, ref_h = synth(B, T, rank)
    mix_h, ref_h = mix_h.pin_memory(), ref_h.pin_memory()
    mix, ref = mix_h.to(dev), ref_h.to(dev)
    rl = torch.tensor(float(T))
    L = P.lib()

    def step_resident():
        return model(mix, ref, rl)

    out_h = torch.empty((B, T), dtype=torch.float32).pin_memory()

    def step_e2e():
        m = mix_h.to(dev, non_blocking=True)
        r = ref_h.to(dev, non_blocking=True)
        est, _ = model(m, r, rl)
        out_h.copy_(est, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller has the separated audio on the host

    def timed(fn, steps, warmup):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = L.launches
            ev0.record()
            for _ in range(steps):
                fn()
            ev1.record()
            barrier()
            ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), L.launches - n0

    with ClockSampler(local) as clk:
        ms_total, launches = timed(step_resident, args.steps, args.warmup)
    clocks = clk.summary()
    ms_e2e, _ = timed(step_e2e, max(1, args.steps), 1)

    audio_per_step = world * B * T / SR
    value = audio_per_step * args.steps / (ms_total / 1e3)
    e2e_value = audio_per_step * max(1, args.steps) / (ms_e2e / 1e3)

    # --- per-kernel pass (separate from the timed region): CUDA events around every launch of the dominant kernel
    dominant = 'dprnn_lstm_layer_bf16' if args.precision == 'bf16' else 'dprnn_lstm_recurrence_f32'
    names = list(L.protos.keys())
    L.timing = {n: [] for n in names}
    with torch.no_grad():
        step_resident()
    torch.cuda.synchronize()
    per_kernel = {}
    for n, evs in L.timing.items():
        if evs:
            d = [a.elapsed_time(b) for a, b in evs]
            per_kernel[n] = {'launches': len(d), 'ms_total': sum(d), 'ms_avg': sum(d) / len(d)}
    L.timing = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)     # kernel timed inside a long step -> sustained figure
    roofline = None
    if dominant in per_kernel:
        k = per_kernel[dominant]
        # one launch = one RNN layer (both directions); the fused bf16 kernel also does the input projection
        flop_per_launch = (RECUR_FLOP_PER_UTT + (PROJ_FLOP_PER_UTT if args.precision == 'bf16' else 0)) * B / 12
        achieved = flop_per_launch / (k['ms_avg'] * 1e-3) / 1e12
        roofline = {'kernel': dominant, 'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                    'frac': achieved / peak_tf, 'traffic': None,
                    'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1400 (B200_PROFILING.md)',
                    'share_of_step': k['ms_total'] / (ms_total / args.steps),
                    'note': ('algorithmic flops = [x_t|h_{t-1}] [W_ih|W_hh]^T, 2*256*512 per chunk position and direction'
                             if args.precision == 'bf16' else
                             'algorithmic flops = h W_hh^T only (2*128*512 per position and direction); the fp32 mode '
                             'runs this on CUDA cores, so its fraction of the bf16 tensor peak is small by construction')}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        kind, cores, times = time_cpu(args.cpu_batch, T, 2, 1)
        v = args.cpu_batch * T / SR * len(times) / sum(times)
        cpu = {'value': v, 'unit': 'audio-s/s', 'cores': cores, 'kind': kind,
               'sample': f'{args.cpu_batch} x {T / SR:g}-s utterance(s), 1 warm-up + 2 timed forwards, fp32, eval()'}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'audio-s/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32' if args.precision == 'fp32' else 'bf16 gate/linear contractions, f32 accumulate+state',
            'data': 'synthetic', 'config': workload_config(args), 'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': 2 * B * T * 4,
                    'd2h_bytes_per_step': B * T * 4},
            'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu, 'kernels': per_kernel,
            'build': L.build_info(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('DPRNN_PRECISION', 'bf16'), choices=['fp32', 'bf16'])
    ap.add_argument('--streams', type=int, default=int(os.environ.get('DPRNN_STREAMS', '1')),
                    help='concurrent CUDA streams the batch is split over inside one forward')
    ap.add_argument('--fused-tail', type=int, default=None, help='1: Linear+norm+residual as one persistent kernel')
    ap.add_argument('--batch', type=int, default=64, help='utterances per GPU (cfg 2: 64)')
    ap.add_argument('--samples', type=int, default=24000, help='samples per utterance (3 s @ 8 kHz)')
    ap.add_argument('--cpu-batch', type=int, default=1, help='utterances per CPU-baseline forward')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
