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
# workloads
# ----------------------------------------------------------------------------------------------------------
NCU_FULL_CAPTURE = 'r2_top_ncu_full.txt'


def ncu_traffic_per_launch(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed `ncu --set full` capture
    (profiles/r2_top_ncu_full.txt; B = 64 cfg-2 shape, fp16 mode) - None if the file or the kernel is missing."""
    try:
        text = open(os.path.join(ROOT, 'profiles', NCU_FULL_CAPTURE)).read()
    except OSError:
        return None
    unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    vals = []
    for block in text.split('Kernel Name = ')[1:]:
        if kernel_substr not in block.splitlines()[0]:
            continue
        tot = 0.0
        for ln in block.splitlines():
            if ln.startswith('dram__bytes_read.sum =') or ln.startswith('dram__bytes_write.sum ='):
                _, rhs = ln.split('=')
                v, u = rhs.split()
                tot += float(v) * unit.get(u, 1.0)
        vals.append(tot)
    return sum(vals) / len(vals) if vals else None


def load_test_set_lengths():
    """(mixture, reference) sample counts of the reference's 3 000 full-length test utterances
    (tests/golden/test_set_lengths.txt, extracted from datasets/tss/test_set.pkl)."""
    rows = []
    with open(os.path.join(ROOT, 'tests', 'golden', 'test_set_lengths.txt')) as f:
        for ln in f:
            if ln.strip() and not ln.startswith('#'):
                a, b = ln.split()
                rows.append((int(a), int(b)))
    return rows


def cfg3_buckets(world, bucket=64):
    """Length-sorted buckets of `bucket` utterances, assigned to ranks by the LPT rule on their chunk count
    (SURVEY.md section 8d cfg 3).  Returns per rank a list of buckets, each a list of (T, Tr)."""
    from tss_with_dprnn_b200.sharding import chunk_count, length_buckets, lpt_assign
    rows = load_test_set_lengths()
    buckets = [[rows[i] for i in idx] for idx in length_buckets([t for t, _ in rows], bucket)]
    costs = [sum(chunk_count(t) for t, _ in bk) for bk in buckets]
    return [[buckets[i] for i in idx] for idx in lpt_assign(costs, world)]


def pick_steps(buckets, n):
    """n buckets spread evenly over a rank's share of the length distribution (a bounded, representative sample)."""
    if n >= len(buckets):
        return list(buckets)
    return [buckets[int(round(i * (len(buckets) - 1) / max(1, n - 1)))] for i in range(n)]


KW4 = dict(KW, embeddings_size=256, fusion_type='att')
KW5 = dict(KW, fusion_type='film')


def model_for(workload):
    if workload == 'cfg1':
        return 'DPRNNTasNet', {k: v for k, v in KW.items() if k != 'fusion_type'}
    if workload == 'cfg3':
        return 'DPRNNSpeIRATasNet', KW
    if workload == 'cfg4':
        return 'DPRNNRawNetTasNet', KW4
    if workload == 'cfg5':
        return 'DPRNNSpeTasNet', KW5
    return 'DPRNNSpeTasNet', KW


def workload_config(args):
    if args.workload == 'cfg3':
        return {'workload': 'cfg3: DPRNN-Spe-IRA (cat) TSS inference on the full-length test-set length distribution '
                            f'(24000..111920 samples), length-sorted buckets of {args.batch} utterances packed as ragged '
                            'batches, buckets LPT-assigned to ranks; one step = one bucket',
                'batch_per_gpu': args.batch, 'precision': args.precision, 'streams': args.streams,
                'l2': 'no flush needed: each step streams >10 GB of intermediates through a 126 MB L2',
                'parallelism': f'utterance sharding x{args.gpus}, no collective'}
    if args.workload == 'cfg1':
        return {'workload': f'cfg1: DPRNN-TasNet BSS forward, 2 speakers, one {args.samples / SR:g}-s 8 kHz mixture, batch '
                            f'{args.batch} (the reference runs this configuration on the CPU: see cpu_baseline)',
                'batch_per_gpu': args.batch, 'samples': args.samples, 'precision': args.precision, 'streams': 1,
                'l2': 'L2 flushed between timed steps (256 MB write): one utterance fits the 126 MB L2',
                'parallelism': f'utterance sharding x{args.gpus}, no collective'}
    if args.workload == 'cfg4':
        return {'workload': f'cfg4: DPRNN-RawNet3 (attention fusion) TSS inference, {args.samples / SR:g}-s mix @ 8 kHz + '
                            f'{args.samples / SR:g}-s raw reference @ 16 kHz, batch {args.batch} per GPU (128 across 8), full '
                            'depth; RawNet3 = stage-1 GPU library restatement, masker/fusion/decoder = CUDA path',
                'batch_per_gpu': args.batch, 'samples': args.samples, 'precision': args.precision,
                'streams': args.streams,
                'l2': 'no flush needed: each step streams >2 GB of intermediates through a 126 MB L2',
                'parallelism': f'utterance sharding x{args.gpus}, no collective'}
    if args.workload == 'cfg5':
        return {'workload': f'cfg5: DPRNN-Spe (FiLM) training step, {args.samples / SR:g}-s crops @ 8 kHz, batch {args.batch} '
                            'per GPU, full depth; one step = zero_grad + forward + (neg SI-SDR + 0.5 CE) + backward + '
                            'all-reduce(mean) of the flat gradient buffer + clip_grad_norm_(5) + Adam(5e-4, wd 1e-5)',
                'batch_per_gpu': args.batch, 'samples': args.samples,
                'precision': ('fp32 (CUDA cores, exact)' if args.precision == 'fp32' else
                              'tensor cores: bf16 LSTM forward + BPTT contractions, TF32 Linear / dX / weight-gradient contractions; fp32 '
                              'accumulation, cell state and element-wise math; speaker encoder exact fp32'), 'streams': 1,
                'samples_per_step': args.batch * args.gpus,
                'l2': 'no flush needed: each step streams >100 GB of saved activations through a 126 MB L2',
                'parallelism': f'data parallel x{args.gpus}, one NCCL all-reduce of 16 MB per step'}
    return {'workload': f'cfg2: DPRNN-Spe (cat) TSS inference, {args.samples / SR:g}-s mix + reference @ 8 kHz, '
                        f'batch {args.batch} per GPU, full depth (6 blocks)',
            'batch_per_gpu': args.batch, 'samples': args.samples, 'precision': args.precision,
            'streams': args.streams,
            'l2': 'no flush needed: each step streams >10 GB of intermediates through a 126 MB L2',
            'parallelism': f'utterance sharding x{args.gpus}, no collective'}


# ----------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own classes (baseline/_ref, installed unmodified) when present,
# else the oracle port (torch CPU restatement calling the same ATen LSTM the reference calls)
# ----------------------------------------------------------------------------------------------------------
def cpu_forward_fn(workload):
    ref_root = os.path.join(ROOT, 'baseline', '_ref')
    torch.manual_seed(0)
    cls, kw = model_for(workload)
    if workload != 'cfg4' and os.path.isdir(os.path.join(ref_root, 'src', 'models')):
        sys.path.insert(0, ref_root)
        if cls == 'DPRNNTasNet':
            from src.models.dprnn import DPRNNTasNet as RefModel
            model = RefModel(**kw).eval()
            return 'reference', (lambda mix, ref, rl: model(mix))
        if cls == 'DPRNNSpeIRATasNet':
            from src.models.dprnn_spe_ira import DPRNNSpeIRATasNet as RefModel
        else:
            from src.models.dprnn_spe import DPRNNSpeTasNet as RefModel
        if workload == 'cfg5':
            return 'reference', cpu_train_step(RefModel(**kw).train())
        model = RefModel(**kw).eval()
        return 'reference', (lambda mix, ref, rl: model(mix, ref, rl)[0])
    from oracle import dprnn_oracle as O
    import tss_with_dprnn_b200 as P
    sd = getattr(P, cls)(**kw).state_dict()
    if workload == 'cfg4':      # the reference's RawNet3 needs asteroid_filterbanks (absent): always the oracle port
        from oracle import rawnet_oracle as RO
        cfg4 = O.Config(fusion_type='att')
        return 'port', (lambda mix, ref, rl: RO.rawnet_tasnet_forward(mix, ref, sd, cfg4)[0])
    if workload == 'cfg1':
        return 'port', (lambda mix, ref, rl: O.tasnet_forward(mix, sd, O.Config()))
    if workload == 'cfg5':      # the port of the model classes is the package's own torch.nn containers + the oracle
        raise RuntimeError('cfg5 CPU arm needs baseline/_ref (the reference model classes drive torch autograd)')
    cfg = O.Config(fusion_type='cat')
    fwd = O.ira_forward if cls == 'DPRNNSpeIRATasNet' else O.spe_forward
    return 'port', (lambda mix, ref, rl: fwd(mix, ref, rl, sd, cfg)[0])


def cpu_train_step(model):
    """The TrainerSpe iteration on the CPU with the reference's own model class (trainer_spe.py:31-56)."""
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=5e-4, weight_decay=1e-5)
    g = torch.Generator().manual_seed(77)

    def step(mix, ref, rl):
        target = 0.05 * torch.randn(mix.shape, generator=g)
        spk = torch.randint(0, 251, (mix.shape[0],), generator=g)
        with torch.enable_grad():
            opt.zero_grad()
            est, logits = model(mix, ref, rl)
            loss = train_loss_cpu(est, target, logits, spk)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5)
            opt.step()
        return float(loss)
    return step


def cpu_sample(args):
    """(batch, T, Tr, description) of the bounded CPU sample of the workload."""
    if args.workload == 'cfg3':
        rows = sorted(load_test_set_lengths())
        t, tr = rows[len(rows) // 2]
        return 1, t, tr, f'1 utterance of the median test-set length ({t} samples, reference {tr}) per forward'
    if args.workload == 'cfg4':
        return args.cpu_batch, args.samples, 2 * args.samples, \
            f'{args.cpu_batch} x {args.samples / SR:g}-s utterance(s) per forward (bounded sample of the batch-{args.batch} workload)'
    return args.cpu_batch, args.samples, args.samples, \
        f'{args.cpu_batch} x {args.samples / SR:g}-s utterance(s) per forward (bounded sample of the batch-{args.batch} workload)'


def time_cpu(args, steps, warmup):
    kind, fwd = cpu_forward_fn(args.workload)
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    batch, T, Tr, desc = cpu_sample(args)
    g = torch.Generator().manual_seed(1234)
    mix = 0.05 * torch.randn(batch, T, generator=g)
    ref = 0.05 * torch.randn(batch, Tr, generator=g)
    rl = torch.tensor(float(Tr))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fwd(mix, ref, rl)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return kind, cores, times, batch * T / SR, desc


def reference_config(args, desc):
    """The workload is the repo arm's (same `workload` string); every other key says what THIS arm ran: the reference's
    CPU code has one precision (fp32), no streams, and is timed on a bounded sample of the batch (CPU throughput is
    flat in the batch size: cpu_baseline.b8)."""
    cfg = workload_config(args)
    cfg.update({'precision': 'fp32 (the reference has no other)', 'streams': 'n/a (CPU)', 'l2': 'n/a (CPU)',
                'parallelism': f'{os.cpu_count()} host threads, rank 0 only', 'batch_per_forward': cpu_sample(args)[0],
                'sample': desc})
    cfg.pop('batch_per_gpu', None)
    return cfg


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    try:
        kind, cores, times, audio, desc = time_cpu(args, args.steps, args.warmup)
    except Exception as e:           # e.g. the cfg5 arm without baseline/_ref: say so instead of dying
        print(json.dumps({'impl': 'reference', 'unavailable': f'{type(e).__name__}: {e}'[:200]}))
        return
    total = sum(times)
    value = audio * len(times) / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'audio-s/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': reference_config(args, desc),
        'cpu_baseline': {'value': value, 'unit': 'audio-s/s', 'cores': cores, 'kind': kind, 'sample': desc},
        'e2e': {'value': value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
class Cfg2:
    """BASELINE.json configs[1]: equal-length 3-s batch through the reference-facing forward(mix, ref, ref_len)."""

    def __init__(self, args, model, rank, dev):
        B, T = args.batch, args.samples
        self.model, self.B, self.T, self.dev = model, B, T, dev
        mix_h, ref_h = synth(B, T, rank)
        self.mix_h, self.ref_h = mix_h.pin_memory(), ref_h.pin_memory()
        self.mix, self.ref = self.mix_h.to(dev), self.ref_h.to(dev)
        self.rl = torch.tensor(float(T))
        self.out_h = torch.empty((B, T), dtype=torch.float32).pin_memory()
        self.n_steps = 1

    def audio(self, i):
        return self.B * self.T / SR

    def positions(self, i):            # chunk positions (LSTM steps x sequences) of the step, per RNN layer
        return self.B * POS_PER_UTT

    def resident(self, i):
        return self.model(self.mix, self.ref, self.rl)

    def e2e(self, i):
        m = self.mix_h.to(self.dev, non_blocking=True)
        r = self.ref_h.to(self.dev, non_blocking=True)
        est, _ = self.model(m, r, self.rl)
        self.out_h.copy_(est, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller has the separated audio on the host

    def bytes(self, i):
        return 2 * self.B * self.T * 4, self.B * self.T * 4


class Cfg1(Cfg2):
    """BASELINE.json configs[0]: DPRNN-TasNet, one 3-s mixture (what the reference's example / test loop runs, B = 1)."""

    def __init__(self, args, model, rank, dev):
        super().__init__(args, model, rank, dev)
        self.out_h = torch.empty((self.B, 2, self.T), dtype=torch.float32).pin_memory()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2: written between timed steps

    def resident(self, i):
        self.flush.zero_()
        return self.model(self.mix)

    def e2e(self, i):
        self.flush.zero_()
        est = self.model(self.mix_h.to(self.dev, non_blocking=True))
        self.out_h.copy_(est, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def bytes(self, i):
        return self.B * self.T * 4, 2 * self.B * self.T * 4


class Cfg4(Cfg2):
    """BASELINE.json configs[3]: DPRNN-RawNet3 (att), raw 16 kHz reference, forward(mix, ref16k)."""

    def __init__(self, args, model, rank, dev):
        super().__init__(args, model, rank, dev)
        g = torch.Generator().manual_seed(1236 + 1000 * rank)
        self.ref_h = (0.05 * torch.randn(self.B, 2 * self.T, generator=g)).pin_memory()
        self.ref = self.ref_h.to(dev)

    def resident(self, i):
        return self.model(self.mix, self.ref)

    def e2e(self, i):
        m = self.mix_h.to(self.dev, non_blocking=True)
        r = self.ref_h.to(self.dev, non_blocking=True)
        est, _ = self.model(m, r)
        self.out_h.copy_(est, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def bytes(self, i):
        return 3 * self.B * self.T * 4, self.B * self.T * 4


def train_loss_cpu(est, target, logits, spk, gamma=0.5):
    """TrainerSpe's loss (trainer_spe.py:39-43) restated for the CPU arm: asteroid's pairwise_neg_sisdr with one source
    (zero-mean, eps 1e-8) + gamma * CrossEntropy."""
    e = est - est.mean(-1, keepdim=True)
    t = target - target.mean(-1, keepdim=True)
    s = (e * t).sum(-1, keepdim=True) * t / (t.pow(2).sum(-1, keepdim=True) + 1e-8)
    sdr = 10 * torch.log10(s.pow(2).sum(-1) / ((e - s).pow(2).sum(-1) + 1e-8) + 1e-8)
    return (-sdr).mean() + gamma * torch.nn.functional.cross_entropy(logits, spk)


class Cfg5:
    """BASELINE.json configs[4]: the TrainerSpe iteration (trainer_spe.py:27-56) on equal-length 3-s crops."""

    def __init__(self, args, model, rank, dev):
        from tss_with_dprnn_b200.train import SpeTrainStep
        B, T = args.batch, args.samples
        self.B, self.T, self.dev = B, T, dev
        self.stepper = SpeTrainStep(model)
        g = torch.Generator().manual_seed(555 + 1000 * rank)
        self.h = [(0.05 * torch.randn(B, T, generator=g)).pin_memory() for _ in range(3)]      # mix, ref, target
        self.spk_h = torch.randint(0, 251, (B,), generator=g).pin_memory()
        self.d = [x.to(dev) for x in self.h]
        self.spk = self.spk_h.to(dev)
        self.loss_h = torch.empty(3).pin_memory()
        self.n_steps = 1

    def audio(self, i):
        return self.B * self.T / SR

    def positions(self, i):
        return self.B * POS_PER_UTT

    def resident(self, i):
        return self.stepper.step(self.d[0], self.d[1], self.d[2], self.spk, ref_len=self.T)

    def e2e(self, i):
        d = [x.to(self.dev, non_blocking=True) for x in self.h]
        spk = self.spk_h.to(self.dev, non_blocking=True)
        loss3 = self.stepper.step(d[0], d[1], d[2], spk, ref_len=self.T)
        self.loss_h.copy_(loss3, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # trainer_spe.py:44 loss.item()

    def bytes(self, i):
        return 3 * self.B * self.T * 4 + self.B * 8, 12


class Cfg3:
    """BASELINE.json configs[2]: full-length variable-duration utterances as ragged batches (model.forward_ragged)."""

    def __init__(self, args, model, rank, world, dev, n_steps):
        self.model, self.dev = model, dev
        mine = cfg3_buckets(world, args.batch)[rank]
        buckets = mine if n_steps is None else pick_steps(mine, n_steps)       # None: the rank's whole share of the test set
        self.n_steps = len(buckets)
        self.steps = []
        g = torch.Generator().manual_seed(4321 + rank)
        for bk in buckets:
            Ts, Trs = [t for t, _ in bk], [tr for _, tr in bk]
            mix_h = (0.05 * torch.randn(sum(Ts), generator=g)).pin_memory()
            ref_h = (0.05 * torch.randn(sum(Trs), generator=g)).pin_memory()
            self.steps.append(dict(Ts=Ts, Trs=Trs, mix_h=mix_h, ref_h=ref_h, mix=mix_h.to(dev), ref=ref_h.to(dev),
                                   out_h=torch.empty(sum(Ts), dtype=torch.float32).pin_memory()))

    def audio(self, i):
        return sum(self.steps[i % self.n_steps]['Ts']) / SR

    def positions(self, i):
        return 2 * sum(250 * ((t - 1 + 250) // 125 + 1) for t in self.steps[i % self.n_steps]['Ts'])   # two masker passes

    def resident(self, i):
        s = self.steps[i % self.n_steps]
        return self.model.forward_ragged((s['mix'], s['Ts']), (s['ref'], s['Trs']))

    def e2e(self, i):
        s = self.steps[i % self.n_steps]
        m = s['mix_h'].to(self.dev, non_blocking=True)
        r = s['ref_h'].to(self.dev, non_blocking=True)
        est, _ = self.model.forward_ragged((m, s['Ts']), (r, s['Trs']))
        flat = torch.cat(est)
        s['out_h'].copy_(flat, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def bytes(self, i):
        s = self.steps[i % self.n_steps]
        return (sum(s['Ts']) + sum(s['Trs'])) * 4, sum(s['Ts']) * 4


# ----------------------------------------------------------------------------------------------------------
# side measurements of the default (cfg 2) line: other precision modes, the reference's modules on this GPU,
# a short cfg-5 training step (the only path with a collective)
# ----------------------------------------------------------------------------------------------------------
def fixture_error(model, dev):
    """Peak-normalised error of the CUDA path against the reference's own fp32 output: fixture spe_cat_r6_3s
    (tests/golden, written by the live reference for torch.manual_seed(0) weights of exactly the bench model)."""
    import numpy as np
    try:
        z = np.load(os.path.join(ROOT, 'tests', 'golden', 'spe_cat_r6_3s.npz'), allow_pickle=False)
    except OSError:
        return None
    mix, ref = torch.from_numpy(z['mix']).to(dev), torch.from_numpy(z['ref']).to(dev)
    want = torch.from_numpy(z['est']).to(dev).double()
    with torch.no_grad():
        est, _ = model(mix, ref, torch.tensor(float(ref.shape[1])))
    return float((est.double() - want).abs().max() / want.abs().max())


def gpu_reference(args, dev):
    """SURVEY.md 8d 'informative extra row': the UNMODIFIED reference classes (baseline/_ref) on this GPU through stock
    PyTorch (cuDNN LSTM, ATen GroupNorm / unfold / fold), same batch, outside the repo arm's timed region."""
    ref_root = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'src', 'models')):
        return {'unavailable': 'baseline/_ref absent'}
    out = {}
    try:
        sys.path.insert(0, ref_root)
        from src.models.dprnn_spe import DPRNNSpeTasNet as RefModel
        torch.manual_seed(0)
        model = RefModel(**KW).eval().to(dev)
        mix, ref = synth(args.batch, args.samples, 0, dev)
        rl = torch.tensor(float(args.samples), device=dev)      # dprnn_spe.py:159-161 divides a device tensor by it
        old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        for name, tf32 in (('fp32', False), ('tf32_allowed', True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad():
                for _ in range(2):
                    model(mix, ref, rl)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    model(mix, ref, rl)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            out[name] = {'ms_per_step': ms, 'value': args.batch * args.samples / SR / (ms / 1e3), 'unit': 'audio-s/s'}
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        out['what'] = (f'baseline/_ref DPRNNSpeTasNet (cat).cuda().eval(), batch {args.batch} x {args.samples / SR:g} s, torch '
                       f'{torch.__version__} / cuDNN {torch.backends.cudnn.version()}, 2 warm-up + 3 timed forwards, CUDA events')
        del model
    except Exception as e:          # an informative row must not take the bench line down
        out['unavailable'] = f'{type(e).__name__}: {e}'[:200]
    torch.cuda.empty_cache()
    return out


def cfg5_sub(args, P, rank, world, dev, barrier, reduce_timing):
    """A short cfg-5 measurement (DPRNN-Spe FiLM training step, 16 x 3 s per GPU, bf16 tensor-core mode) so that the
    driver's 1 -> 8 GPU runs of the DEFAULT command also exercise the one collective of the project: the NCCL
    all-reduce of the flat gradient buffer.  Ranks are seeded DIFFERENTLY on purpose: the stepper has to broadcast rank
    0's parameters, and replica_param_checksum_spread must come out 0."""
    import torch.distributed as dist
    a5 = argparse.Namespace(batch=16, samples=24000)
    torch.manual_seed(100 + rank)
    model = P.DPRNNSpeTasNet(**KW5).train().to(dev)
    model.precision = 'bf16'
    wl = Cfg5(a5, model, rank, dev)
    wl.stepper.allreduce_events = []
    steps, warm = 3, 2
    for i in range(warm):
        wl.resident(i)
    barrier()
    wl.stepper.allreduce_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        wl.resident(i)
    e1.record()
    barrier()
    ms_local = e0.elapsed_time(e1)
    ms, audio = reduce_timing(ms_local, steps * wl.audio(0), dev)
    ar = [a.elapsed_time(b) for a, b in wl.stepper.allreduce_events]
    chk = wl.stepper.fp.flat.double().sum().reshape(1)
    spread = 0.0
    if world > 1:
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        spread = float(max(float(c) for c in allc) - min(float(c) for c in allc))
    out = {'workload': 'cfg5: DPRNN-Spe (FiLM) training step, 16 x 3 s per GPU, bf16 tensor-core mode; 2 warm-up + 3 timed '
                       'steps; ranks start from different seeds (parameters broadcast from rank 0)',
           'ms_per_step': ms / steps, 'samples_per_s': 16 * world * steps / (ms / 1e3), 'audio_s_per_s': audio / (ms / 1e3),
           'allreduce_ms': (sum(ar) / len(ar)) if ar else 0.0, 'allreduce_bytes': int(wl.stepper.fp.size * 4),
           'replica_param_checksum_spread': spread, 'n_gpus': world}
    del wl, model
    torch.cuda.empty_cache()
    return out


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
    cls, kw = model_for(args.workload)
    model = getattr(P, cls)(**kw).eval().to(dev)      # same seeded weights on every rank (replicated, 16 MB)
    model.precision = args.precision
    model.n_streams = args.streams
    model._engine.lstm_slices = args.lstm_slices
    model._engine.lstm_pairs = args.lstm_pairs
    model._engine.residual_bf16 = bool(args.residual_bf16)
    model._engine.lstm_pingpong = bool(args.lstm_pingpong)
    if args.fused_tail is not None:
        model._engine.fused_tail = bool(args.fused_tail)
    L = P.lib()
    if args.workload == 'cfg3':
        wl = Cfg3(args, model, rank, world, dev, None if args.full else args.steps)
    elif args.workload == 'cfg4':
        wl = Cfg4(args, model, rank, dev)
    elif args.workload == 'cfg1':
        wl = Cfg1(args, model, rank, dev)
    elif args.workload == 'cfg5':
        wl = Cfg5(args, model.train(), rank, dev)
    else:
        wl = Cfg2(args, model, rank, dev)

    def timed(fn, steps, warmup):
        with torch.no_grad():
            for i in range(warmup):
                fn(i)
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]     # end of every step: median / best
            n0 = L.launches
            ev0.record()
            for i in range(steps):
                fn(i)
                marks[i].record()
            ev1.record()
            barrier()
            ms = ev0.elapsed_time(ev1)
            timed.per_step = [a.elapsed_time(b) for a, b in zip([ev0] + marks[:-1], marks)]
        return ms, L.launches - n0

    from tss_with_dprnn_b200.sharding import reduce_timing
    # cfg3 --full: every rank runs ALL the buckets the LPT rule gave it (the whole 3 000-utterance test set across the
    # job), so the per-rank step counts differ and the max-over-ranks time shows the real imbalance
    steps_local = wl.n_steps if (args.workload == 'cfg3' and args.full) else args.steps
    audio_local = sum(wl.audio(i) for i in range(steps_local))
    with ClockSampler(local) as clk:
        # every distinct step is warmed up at least once (ragged layouts, tensor maps, allocator pools)
        ms_local, launches = timed(wl.resident, steps_local, max(args.warmup, wl.n_steps))
    clocks = clk.summary()
    step_ms = sorted(timed.per_step)          # this rank's steps of the timed region (median and best)
    ms_total, audio_steps = reduce_timing(ms_local, audio_local, dev)     # max over ranks of device time, whole-job audio
    ms_local_e2e, _ = timed(wl.e2e, steps_local, max(1, wl.n_steps))
    ms_e2e, _ = reduce_timing(ms_local_e2e, audio_local, dev)
    value = audio_steps / (ms_total / 1e3)
    e2e_value = audio_steps / (ms_e2e / 1e3)
    h2d = sum(wl.bytes(i)[0] for i in range(steps_local)) / steps_local
    d2h = sum(wl.bytes(i)[1] for i in range(steps_local)) / steps_local
    rank_ms = [ms_local]
    if world > 1:
        t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [float(x) for x in allt]

    # --- per-kernel pass (separate from the timed region, one stream so that launches do not overlap): CUDA events
    # around every launch of step 0
    tcmode = args.precision != 'fp32'
    dominant = 'dprnn_lstm_layer_bf16' if tcmode else 'dprnn_lstm_recurrence_f32'
    names = list(L.protos.keys())
    model.n_streams = 1
    # three passes; the one with the smallest total is reported (the first single-stream pass pays for first-use effects -
    # new buffer sizes, lazily loaded kernels - and a shared host can stall any one of them); every pass's dominant-kernel
    # average launch is kept in roofline.passes_top_kernel_ms_avg
    per_kernel, pass_totals, pass_dominant = {}, [], []
    for _ in range(3):
        L.timing = {n: [] for n in names}
        with torch.no_grad():
            wl.resident(0)
        torch.cuda.synchronize()
        pk = {}
        for n, evs in L.timing.items():
            if evs:
                d = [a.elapsed_time(b) for a, b in evs]
                pk[n] = {'launches': len(d), 'ms_total': sum(d), 'ms_avg': sum(d) / len(d)}
        L.timing = None
        tot = sum(v['ms_total'] for v in pk.values())
        pass_totals.append(tot)
        top = max(pk.values(), key=lambda v: v['ms_total']) if pk else None
        pass_dominant.append(top['ms_avg'] if top else None)
        if not per_kernel or tot <= min(pass_totals[:-1]):
            per_kernel = pk
    model.n_streams = args.streams
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)     # kernel timed inside a long step -> sustained figure
    roofline = None
    lstm_names = [n for n in per_kernel if n in (dominant, 'dprnn_lstm_inter_bf16_ragged', 'dprnn_lstm_layer_bf16_pp',
                                                 'dprnn_lstm_layer_bf16_sliced', 'dprnn_lstm_inter_bf16_ragged_pp')]
    if args.workload == 'cfg5':
        lstm_names = [n for n in per_kernel if n in ('dprnn_lstm_recurrence_f32_train', 'dprnn_lstm_bptt_f32', 'dprnn_lstm_layer_bf16_train', 'dprnn_lstm_bptt_tc', 'dprnn_lstm_bptt_tc_bf16out')]
    if lstm_names:
        ms_lstm = sum(per_kernel[n]['ms_total'] for n in lstm_names)
        n_launch = sum(per_kernel[n]['launches'] for n in lstm_names)
        # one launch = one RNN layer (both directions) over every chunk position of the step; the fused bf16 kernel
        # also does the input projection: 2 * 2 * 128 * 512 flop per position and direction (else 2 * 128 * 512)
        flop_per_pos = 2 * (2 if tcmode and args.workload != 'cfg5' else 1) * 2 * 128 * 512
        flop_per_launch = flop_per_pos * wl.positions(0) / (2 if args.workload == 'cfg3' else 1)
        achieved = flop_per_launch * n_launch / (ms_lstm * 1e-3) / 1e12
        roofline = {'kernel': '+'.join(lstm_names), 'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf,
                    'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                    'traffic': (ncu_traffic_per_launch('lstm_tc_sliced' if 'dprnn_lstm_layer_bf16_sliced' in lstm_names
                                                       else 'lstm_tc_pp')
                                if args.workload == 'cfg2' and args.batch == 64 and tcmode else None),
                    'traffic_note': f'DRAM read+write bytes per launch, ncu --set full (profiles/{NCU_FULL_CAPTURE}); '
                                    'algorithmic: read xb 2 x 0.79 GB + write hb 1.59 GB = 3.18 GB',
                    'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1400 (B200_PROFILING.md)',
                    'launch_ms_avg': ms_lstm / n_launch,
                    'share_of_step_single_stream': ms_lstm / sum(k['ms_total'] for k in per_kernel.values()),
                    'note': ('algorithmic flops = [x_t|h_{t-1}] [W_ih|W_hh]^T, 2*256*512 per chunk position and direction; '
                             'launches timed one by one on a single stream'
                             if tcmode and args.workload != 'cfg5' else
                             'algorithmic flops = h W_hh^T only (2*128*512 per position and direction); the fp32 mode '
                             'runs this on CUDA cores, so its fraction of the bf16 tensor peak is small by construction')}

    bptt_name = next((n for n in ('dprnn_lstm_bptt_tc_bf16out', 'dprnn_lstm_bptt_tc', 'dprnn_lstm_bptt_f32') if n in per_kernel),
                     'dprnn_lstm_bptt_f32')
    if args.workload == 'cfg5' and bptt_name in per_kernel:
        # dominant kernel of the training step: the BPTT recurrence.  Per chunk position and direction it streams the saved
        # gate activations (4 H), the cell state (H; c_{t-1} of a step is c_t of the next one) and d h_out (H) and writes
        # d gates (4 H).
        k = per_kernel[bptt_name]
        peak_bw = peaks.get('hbm_gbs', 6400.0)
        H_ = 128
        gate_bytes = 4 * H_ * (4 if bptt_name == 'dprnn_lstm_bptt_f32' else 2)           # saved gates: fp32 / packed bf16
        dg_bytes = 4 * H_ * (2 if bptt_name == 'dprnn_lstm_bptt_tc_bf16out' else 4)      # d gates out: bf16 / fp32
        c_bytes = H_ * 4 * (1 if bptt_name == 'dprnn_lstm_bptt_tc_bf16out' else 2)       # older forms load c_t and c_{t-1}
        bytes_per_launch = wl.positions(0) * 2 * (gate_bytes + c_bytes + H_ * 4 + dg_bytes)
        achieved = bytes_per_launch / (k['ms_avg'] * 1e-3) / 1e9
        roofline = {'kernel': bptt_name, 'bound': 'hbm', 'achieved': achieved, 'peak': peak_bw, 'unit': 'GB/s',
                    'frac': achieved / peak_bw, 'traffic': None,
                    'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6400 (B200_PROFILING.md)',
                    'launch_ms_avg': k['ms_avg'],
                    'share_of_step_single_stream': k['ms_total'] / sum(v['ms_total'] for v in per_kernel.values()),
                    'bytes_per_position_and_direction': gate_bytes + c_bytes + H_ * 4 + dg_bytes,
                    'note': 'dominant kernel of the training step (BPTT).  Algorithmic bytes per chunk position and direction: '
                            '4 H saved gate activations (bf16 in tensor-core mode), c_t and d h_out (fp32) in; d gates out (4 H; '
                            'bf16 when the weight-gradient and d x kernels take bf16 operands).  d h_{t-1} = d gates_t W_hh on '
                            'tcgen05 (CTA pair, 64 rows per CTA at 16 utterances per GPU: 100-128 of 148 SMs hold a CTA); the '
                            'element-wise cell backward between the MMAs is issue- and latency-bound (ncu: 47 % issue slots, top '
                            'stalls = first use of the streamed loads and the wait for the step\'s MMA).  The launch is timed while '
                            'the weight-gradient kernels of the previous half-block run on the side stream and share the HBM '
                            'bandwidth with it: DESIGN.md section 5.2'}

    if roofline is not None:      # the three per-kernel passes: sum of all launches, and the top kernel's average launch, per pass
        roofline['passes_total_ms'] = pass_totals
        roofline['passes_top_kernel_ms_avg'] = pass_dominant
        roofline['pass_reported'] = 'the pass with the smallest total (index %d)' % pass_totals.index(min(pass_totals))
    modes = gpu_ref = cfg5 = None
    if args.workload == 'cfg2':
        if args.modes and world == 1:
            # the other precision modes on the same batch: throughput, and the error of each against the reference's own
            # fp32 output (fixture of the same seeded weights); tol_1e-3 names the fastest mode inside north_star's tolerance
            modes = {args.precision: {'value': value, 'ms_per_step': ms_total / args.steps,
                                      'err': fixture_error(model, dev)}}
            for mode in ('fp16', 'bf16', 'fp32'):
                if mode in modes:
                    continue
                model.precision = mode
                k, w = (1, 1) if mode == 'fp32' else (3, 3)
                ms_m, _ = timed(wl.resident, k, w)
                modes[mode] = {'value': k * wl.audio(0) / (ms_m / 1e3), 'ms_per_step': ms_m / k, 'err': fixture_error(model, dev)}
            model.precision = args.precision
            ok = [m for m, v in modes.items() if v['err'] is not None and v['err'] <= 1e-3]
            modes['tol_1e-3'] = dict(modes[max(ok, key=lambda m: modes[m]['value'])], mode=max(ok, key=lambda m: modes[m]['value'])) if ok else None
            modes['err_metric'] = ('max|est - ref| / max|ref| against tests/golden/spe_cat_r6_3s.npz (the live reference, fp32, '
                                   'same seeded weights), B = 1')
        model._engine.invalidate()          # drop the captured graphs (they pin their forward's buffers)
        torch.cuda.empty_cache()
        if args.gpu_reference and world == 1:
            gpu_ref = gpu_reference(args, dev)
        if args.cfg5:
            del wl
            torch.cuda.empty_cache()
            cfg5 = cfg5_sub(args, P, rank, world, dev, barrier, reduce_timing)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        # SURVEY.md section 8d protocol: 1 warm-up + 3 timed, best-of and mean, at B = 1 and (cfg 2) B = 8
        kind, cores, times, audio, desc = time_cpu(args, 3, 1)
        cpu = {'value': audio * len(times) / sum(times), 'best': audio / min(times), 'unit': 'audio-s/s', 'cores': cores,
               'kind': kind,
               'sample': desc + (', 1 warm-up + 3 timed training iterations, fp32, train()' if args.workload == 'cfg5'
                                  else ', 1 warm-up + 3 timed forwards, fp32, eval()')}
        if args.workload == 'cfg2' and args.cpu_b8:
            a8 = argparse.Namespace(**vars(args))
            a8.cpu_batch = 8
            _, _, t8, audio8, desc8 = time_cpu(a8, 3, 1)
            cpu['b8'] = {'value': audio8 * len(t8) / sum(t8), 'best': audio8 / min(t8), 'sample': desc8 + ', 1 warm-up + 3 timed'}

    replica_spread = None
    if args.workload == 'cfg5':
        # data-parallel replicas must hold identical parameters after the all-reduced updates
        chk = wl.stepper.fp.flat.double().sum().reshape(1)
        if world > 1:
            allc = [torch.zeros_like(chk) for _ in range(world)]
            dist.all_gather(allc, chk)
            replica_spread = float(max(float(c) for c in allc) - min(float(c) for c in allc))
        else:
            replica_spread = 0.0
    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'audio-s/s', 'n_gpus': world, 'steps': steps_local,
            'warmup': args.warmup, 'ms_per_step': ms_total / steps_local, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32' if args.precision == 'fp32' else 'bf16 LSTM gate contractions (fwd + BPTT), tf32 linear / weight-gradient contractions, f32 accumulate+state' if args.workload == 'cfg5' else f'{args.precision} gate/linear contractions (tcgen05), tf32 1x1 convs, f32 accumulate+state',
            'data': 'synthetic', 'config': workload_config(args), 'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
            'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu, 'kernels': per_kernel,
            'ranks': {'ms': rank_ms, 'imbalance_max_over_mean': max(rank_ms) / (sum(rank_ms) / len(rank_ms)),
                      'steps_rank0': steps_local},
            'build': L.build_info(),
            # rank 0's own steps inside the timed region (events at the end of every step on the launching stream)
            'step_ms': {'median': step_ms[len(step_ms) // 2], 'best': step_ms[0], 'worst': step_ms[-1]},
        }
        if modes is not None:
            line['modes'] = modes
        if gpu_ref is not None:
            line['gpu_reference'] = gpu_ref
        if cfg5 is not None:
            line['cfg5'] = cfg5
        if args.workload == 'cfg5':
            line['train'] = {'samples_per_s': args.batch * world * args.steps / (ms_total / 1e3),
                             'replica_param_checksum_spread': replica_spread,
                             'grad_allreduce_bytes': int(wl.stepper.fp.size * 4)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=['cfg1', 'cfg2', 'cfg3', 'cfg4', 'cfg5'],
                    help='cfg1: DPRNN-TasNet B=1; cfg2 (headline): DPRNN-Spe cat, 3-s batch 64; cfg3: DPRNN-Spe-IRA on the ragged test-set lengths; '
                         'cfg4: DPRNN-RawNet3 att, batch 16 per GPU; cfg5: DPRNN-Spe FiLM training step, batch 16 per GPU')
    ap.add_argument('--precision', default=None, choices=['fp32', 'bf16', 'fp16'],
                    help="default: fp16 (tensor-core mode inside north_star's 1e-3 tolerance); cfg5: bf16")
    ap.add_argument('--streams', type=int, default=int(os.environ.get('DPRNN_STREAMS', '3')),
                    help='concurrent CUDA streams the batch is split over inside one forward (cfg2)')
    ap.add_argument('--lstm-slices', type=int, default=int(os.environ.get('DPRNN_LSTM_SLICES', '1')),
                    help='time slices per LSTM job of the persistent kernel (1 = one job per CTA pair)')
    ap.add_argument('--lstm-pingpong', type=int, default=int(os.environ.get('DPRNN_LSTM_PINGPONG', '1')),
                    help='1 (default): half-job ping-pong LSTM kernel; 0: one job per CTA pair')
    ap.add_argument('--residual-bf16', type=int, default=1,
                    help='tensor-core modes: 1 (default) residual stream in 16 bits only; 0: fp32 master copy of the residual stream')
    ap.add_argument('--lstm-pairs', type=int, default=int(os.environ.get('DPRNN_LSTM_PAIRS', '0')),
                    help='cap on the resident CTA pairs of the persistent LSTM kernel (0 = all)')
    ap.add_argument('--fused-tail', type=int, default=None, help='1: Linear+norm+residual as one persistent kernel')
    ap.add_argument('--batch', type=int, default=None, help='utterances per GPU and step (cfg 2: 64; cfg 3: bucket size)')
    ap.add_argument('--samples', type=int, default=24000, help='samples per utterance (3 s @ 8 kHz)')
    ap.add_argument('--cpu-batch', type=int, default=1, help='utterances per CPU-baseline forward')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--full', action='store_true',
                    help='cfg3: every rank runs its whole LPT share of the 3 000-utterance test set (steps = its bucket count)')
    ap.add_argument('--cpu-b8', type=int, default=1, help='cfg2: also time the CPU baseline at B = 8 (SURVEY.md 8d protocol)')
    ap.add_argument('--modes', type=int, default=1,
                    help='cfg2: also measure the other precision modes (throughput + error against the reference fixture)')
    ap.add_argument('--gpu-reference', type=int, default=1,
                    help='cfg2, 1 GPU: also time the unmodified reference classes on this GPU through stock PyTorch / cuDNN')
    ap.add_argument('--cfg5', type=int, default=1,
                    help='cfg2: append a short cfg-5 (training step, all-reduce) sub-measurement to the line')
    args = ap.parse_args()
    if args.batch is None:
        args.batch = {'cfg1': 1, 'cfg4': 16, 'cfg5': 16}.get(args.workload, 64)
    if args.precision is None:
        args.precision = os.environ.get('DPRNN_PRECISION', 'bf16' if args.workload == 'cfg5' else 'fp16')
    if args.workload in ('cfg1', 'cfg5'):
        args.streams = 1
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
