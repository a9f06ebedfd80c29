"""Headline benchmark: captions/sec, beam 5, 32 tokens (BASELINE.json), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg5|cfg4|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one pass of the caption path over one batch of synthetic images: ResNet-50 encoder (+ label encoder)
-> decoder -> stochastic beam search (beam 5, top-k 50, max_len 32) -> token ids.  Default workload at every N is
BASELINE.json configs[1]: CaptioningLSTMWithLabels (emb 512 / hidden 512 / 3 layers, V = 36 541), 512 images per
GPU (weak scaling: images are sharded by global index, no collective in the loop, ONE all-gather of ids at the end
of each step).  `value` = images all ranks processed / max-over-ranks device time, inputs resident in HBM; `e2e`
= same through model.generate() with pinned HOST images copied in and ids copied out every step.

--impl reference times the CPU restatement of the reference (oracle/, kind "port": /root/reference is pure
Python + torch and does not exist on the GPU box) on a bounded sample of the same workload, all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

V = 36541
WORKLOADS = {
    # name: (kind, per-GPU batch, beam, top_k, description)
    'cfg2': ('lstm_labels', 512, 5, 50, 'CaptioningLSTMWithLabels beam5 top_k50 max_len32, 512 images/GPU, V=36541'),
    'cfg5': ('xfmr', 8192, 5, 50, 'CaptioningTransformer beam5 top_k50 max_len32, 8192 images/GPU, V=36541'),
    'cfg4': ('xfmr', 4096, 1, 50, 'CaptioningTransformer top_k50 sampling max_len32, 4096 images/GPU, V=36541'),
    'cfg1': ('lstm', 8, 1, 1, 'CaptioningLSTM greedy max_len32, 8 images/GPU, V=36541'),
    # teacher-forced perplexity eval (value = sequences/s): encoder + decoder over 32 positions + fused log-softmax
    'cfg3': ('xfmr_base', 2048, 0, 0, 'CaptioningTransformerBase teacher-forced perplexity, 2048 x 32 tokens/GPU, V=36541'),
}
MAX_LEN = 32
METRIC = 'captions/sec (beam 5, 32 tok)'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '25', '-i', str(self.index)],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


def build_model(kind, precision):
    from deephumor_b200 import models
    from deephumor_b200.utils import synth_weights
    cls = {'lstm': models.CaptioningLSTM, 'lstm_labels': models.CaptioningLSTMWithLabels,
           'xfmr_base': models.CaptioningTransformerBase, 'xfmr': models.CaptioningTransformer}[kind]
    hp = synth_weights.default_hp(kind, V)
    if kind == 'lstm':            # BASELINE.json configs[0]: 1-layer LSTMDecoder, emb 256
        hp.update(emb_dim=256, num_layers=1)
    sd = synth_weights.make_state_dict(kind, hp, seed=0)
    m = cls(**hp)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval().set_precision(precision), hp, sd


def algorithmic_flops(kind, beam, n_img):
    """2*MAC per caption (encoder 8.174 G + heads + decoder row-steps)."""
    enc = 8.174e9 + 2 * 2048 * 512
    if kind in ('lstm', 'lstm_labels'):
        row_step = 3 * 2 * 4 * 512 * 1024 + 2 * 512 * V
        steps = 1 + (MAX_LEN - 1) * beam
        return n_img * (enc + steps * row_step)
    per_layer = 2.097e6 + 4.194e6 + (1.049e6 if kind == 'xfmr' else 0) + 0.17e6
    row_step = 3 * per_layer + 2 * 512 * V
    steps = 1 + MAX_LEN * beam
    extra = (49 * 2 * 2048 * 512 + 154e6) if kind == 'xfmr' else 0
    return n_img * (enc + extra + steps * row_step)


def traffic_bytes(args, rows):
    """DRAM bytes per launch of the roofline kernel from the committed `ncu --set full` capture (same shape only)."""
    p = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if args.precision != 'bf16' or rows != 2560 or not os.path.exists(p):
        return None
    return json.load(open(p)).get('dram_bytes_per_launch')


def run_ours(args):
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from deephumor_b200 import _lib
    from deephumor_b200.runtime import ops, shard
    from deephumor_b200.utils import synth
    kind, batch, beam, top_k, desc = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    model, hp, sd = build_model(kind, args.precision)
    dev = torch.device('cuda', local)
    first = rank * batch
    images = torch.empty(batch, 3, 224, 224, device=dev)
    ops.synth_images(images, 0, first)
    labels = synth.labels(0, first, batch, V).to(dev) if kind == 'lstm_labels' else None
    host_images = torch.empty(batch, 3, 224, 224, pin_memory=True)
    host_images.copy_(images)
    host_labels = labels.cpu().pin_memory() if labels is not None else None
    gen_kw = dict(max_len=MAX_LEN, temperature=1.0, beam_size=beam, top_k=top_k, noise='injected', seed=1234,
                  image_base=first)

    captions = cap_lens = None
    if args.workload == 'cfg3':
        captions, cap_lens = synth.captions(0, first, batch, V, width=MAX_LEN, min_len=8)
        captions, cap_lens = captions.to(dev), cap_lens.to(dev)

    def step(img, lab):
        if captions is not None:                     # config 3: one scalar leaves the device
            with torch.no_grad():
                pp = model.perplexity(img, captions, cap_lens)
            return pp.view(1, 1), pp.view(1)
        with torch.no_grad():
            out = model.generate(img, lab, **gen_kw) if lab is not None else model.generate(img, **gen_kw)
        ids, lens = out if isinstance(out, tuple) else (out.view(1, -1), torch.tensor([out.numel()], device=dev))
        # the path's only collective: one all-gather of ids (+lengths) per step; identity at N=1
        return shard.gather_captions(ids, lens, total=batch * world)

    def step_e2e():
        # the user-facing call with HOST buffers: generate() streams the pinned images to the device in chunks
        # (copy of chunk k+1 overlapping the trunk of chunk k) and the ids / lengths are read back
        lab = host_labels.to(dev, non_blocking=True) if host_labels is not None else None
        ids, lens = step(host_images, lab)
        return ids.cpu(), (lens.cpu() if lens is not None else None)

    def timed(fn, steps, profile=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.LIB.launches
        if profile:
            ops.PROFILE.start()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        prof = ops.PROFILE.stop() if profile else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item()), _lib.LIB.launches - l0, prof

    if args.profile_mode:       # under `ncu --profile-from-start off`: warm-up, then ONE eagerly launched step
        step(images, labels)
        step(images, labels)
        torch.cuda.synchronize()
        ops.USE_GRAPHS = False
        torch.cuda.profiler.start()
        step(images, labels)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    for _ in range(max(args.warmup, 3)):
        step(images, labels)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, _ = timed(lambda: step(images, labels), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # per-stage / per-kernel CUDA-event ranges: a second pass over the same steps, launched eagerly because the
    # timed pass replays the decode loop as one CUDA graph (events cannot be recorded inside a replay)
    _, launches_eager, prof = timed(lambda: step(images, labels), args.steps, profile=True)
    launches = max(launches, launches_eager)
    step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    # extension: raw uint8 pixels from the host, ToTensor + Normalize fused into the stem kernel
    ms_e2e_u8 = None
    if args.precision == 'bf16' and captions is None:
        host_u8 = torch.randint(0, 256, (batch, 3, 224, 224), dtype=torch.uint8).pin_memory()

        def step_e2e_u8():
            lab = host_labels.to(dev, non_blocking=True) if host_labels is not None else None
            ids, lens = step(host_u8, lab)
            return ids.cpu(), lens.cpu()
        step_e2e_u8()
        ms_e2e_u8, _, _ = timed(step_e2e_u8, args.steps)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    total = batch * world
    value = total * args.steps / (ms / 1e3)
    e2e = total * args.steps / (ms_e2e / 1e3)
    pk, pk_src = peaks()
    line = {
        'metric': METRIC if captions is None else 'sequences/sec (teacher-forced perplexity, 32 tok)',
        'value': round(value, 2), 'unit': 'captions/s' if captions is None else 'sequences/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': round(ms / args.steps, 3), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': args.precision if args.precision != 'fp32' else 'f32',
        'data': 'synthetic (hash-generated 224x224 images by global index, random-init weights seed 0)',
        'config': {'workload': f'{args.workload}: {desc}', 'images_per_gpu': batch, 'global_batch': total,
                   'max_len': MAX_LEN, 'beam_size': beam, 'top_k': top_k, 'noise': 'injected',
                   'l2': 'inputs larger than L2 (308 MB fp32 images per GPU per step)',
                   'parallelism': f'dp{world} (images sharded by global index; one all-gather of ids per step)'},
        'decode_tokens_per_s': round(value * MAX_LEN, 1),
        'e2e': {'value': round(e2e, 2), 'unit': 'captions/s', 'ms_per_step': round(ms_e2e / args.steps, 3),
                'h2d_bytes_per_step': int(host_images.numel() * 4 + (host_labels.numel() * 8 if host_labels is not None else 0)),
                'd2h_bytes_per_step': int(batch * MAX_LEN * 8 + batch * 8)},
        'gpu_launches': int(launches), 'clocks': clocks,
    }
    if ms_e2e_u8 is not None:
        line['e2e_uint8'] = {'value': round(total * args.steps / (ms_e2e_u8 / 1e3), 2), 'unit': 'captions/s',
                             'ms_per_step': round(ms_e2e_u8 / args.steps, 3), 'h2d_bytes_per_step': int(batch * 3 * 224 * 224),
                             'note': 'host uint8 pixels, preprocessing fused into the stem kernel (not the headline e2e)'}
    # ---- roofline of the dominant kernel (the vocab-projection contraction)
    flops_step = algorithmic_flops(kind, beam, batch)
    sustained = pk.get('bf16_tflops_sustained', pk.get('bf16_tflops'))
    line['step_tflops'] = round(flops_step / (ms / args.steps / 1e3) / 1e12, 2)
    if prof and prof.get('vocab_gemm'):
        n, tot_ms, flops, _ = prof['vocab_gemm']
        ach = flops / (tot_ms / 1e3) / 1e12
        tensor_path = args.precision == 'bf16'
        line['roofline'] = {'kernel': 'gemm_tc_kernel, candidate-compaction epilogue (vocab projection [rows,512]x[512,36541], '
                                      'pass 2 of the fused selection; logits never stored)' if tensor_path
                            else 'igemm_f32_kernel (fp32 check mode FFMA)',
                            'bound': 'tensor', 'achieved': round(ach, 2), 'peak': sustained, 'unit': 'TFLOP/s',
                            'frac': round(ach / sustained, 4), 'traffic': traffic_bytes(args, batch * beam), 'peak_source': f'{pk_src}, sustained bf16',
                            'launches': n, 'avg_ms': round(tot_ms / n, 4),
                            'share_of_step': round(tot_ms / args.steps / (ms / args.steps), 4)}
        line['roofline']['note'] = ('kernel timed with CUDA events in an eager pass of the same steps right after the '
                                    'timed pass (which replays the decode loop as a CUDA graph)')
        line['stages_ms_per_step'] = {k: round(v[1] / args.steps, 3) for k, v in prof.items()}
        # per-stage tensor-pipe fraction (algorithmic FLOPs of the stage / its CUDA-event time / measured sustained peak)
        line['stages_tensor_frac'] = {k: round(v[2] / (v[1] / 1e3) / 1e12 / sustained, 3)
                                      for k, v in prof.items() if v[2] > 0 and v[1] > 0}
    if not args.no_cpu_baseline and world == 1:          # reported at N = 1 only
        line['cpu_baseline'] = cpu_baseline(kind, hp, sd, beam, top_k, args.cpu_images)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(kind, hp, sd, beam, top_k, n_img):
    """The oracle port (CPU restatement of the reference, per-image loop like the reference) on a bounded sample."""
    from deephumor_b200.utils import synth
    from oracle import model as omodel, noise as onoise
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    imgs = synth.images(0, 0, n_img)
    labs = synth.labels(0, 0, n_img, V) if kind == 'lstm_labels' else None
    if beam == 0:
        # teacher-forced workload (config 3): the reference's forward + experiments/metrics.perplexity on a batch
        caps, lens = synth.captions(0, 0, n_img, V, width=MAX_LEN, min_len=8)
        with torch.no_grad():
            run = lambda n: omodel.perplexity(omodel.forward(kind, sd, hp, imgs[:n], caps[:n, :-1]), caps[:n], lens[:n])
            run(1)                                                                                       # warm-up
            t0 = time.time()
            run(n_img)
            dt = time.time() - t0
        return {'value': round(n_img / dt, 3), 'unit': 'sequences/s', 'cores': cores, 'kind': 'port',
                'sample': f'{n_img} sequences of the same workload in one batch, torch CPU fp32, {cores} threads, {dt:.1f} s'}
    kw = dict(max_len=MAX_LEN, beam_size=beam, top_k=top_k, temperature=1.0, noise=onoise.Noise('injected', 1234),
              faithful_cost=True)
    with torch.no_grad():
        omodel.generate_batch(kind, sd, hp, imgs[:1], None if labs is None else labs[:1], **kw)      # warm-up
        t0 = time.time()
        omodel.generate_batch(kind, sd, hp, imgs, labs, **kw)
        dt = time.time() - t0
    return {'value': round(n_img / dt, 3), 'unit': 'captions/s', 'cores': cores, 'kind': 'port',
            'sample': f'{n_img} images of the same workload, one at a time (the reference generate() is batch-1), '
                      f'torch CPU fp32, {cores} threads, {dt:.1f} s'}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from deephumor_b200.utils import synth_weights
    kind, batch, beam, top_k, desc = WORKLOADS[args.workload]
    hp = synth_weights.default_hp(kind, V)
    sd = synth_weights.make_state_dict(kind, hp, seed=0)
    n = min(args.cpu_images, 32)          # ~5 s of host work per step keeps K + W steps within a few minutes
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(kind, hp, sd, beam, top_k, n)
        if i >= args.warmup:
            vals.append(cb['value'])
    value = sum(vals) / len(vals)
    cb['value'] = round(value, 3)
    world = int(os.environ.get('WORLD_SIZE', 1))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': round(value, 3), 'unit': 'captions/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(n / value * 1e3, 1), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: {desc}', 'sample_images_per_step': n, 'max_len': MAX_LEN,
                   'beam_size': beam, 'top_k': top_k},
        'cpu_baseline': cb,
        'e2e': {'value': round(value, 3), 'unit': 'captions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--batch', type=int, default=0, help='override images per GPU')
    ap.add_argument('--cpu-images', type=int, default=64, help='bounded CPU sample size (about 10 s of host work)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-mode', action='store_true', help='1 warm-up + 1 step only (for ncu launch lists)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
