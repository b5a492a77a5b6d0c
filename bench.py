#!/usr/bin/env python3
"""Benchmark of the MeTRo inference hot path on B200 (BASELINE.json metric: 256x256 crops/sec;
soft-argmax HBM GB/s vs measured peak).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config B|D|A|C|E]

A step = one pass of the hot path (crops -> [N,J,3] mm) over one synthetic batch.  At N=1 the
workload is BASELINE config B (ResNet-50, stride 16, 17 joints, batch 256); with torchrun every rank
runs the same per-GPU batch on its own shard (weak scaling) and the step ends with the all-gather of
the results.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

from metro_pose3d_b200.joints import export_permutation, model_joint_info
from metro_pose3d_b200.spec import CONFIGS, NetSpec
from metro_pose3d_b200.weights import synth_head, synth_images, synth_weights

METRIC = '256x256 crops/sec'
UNIT = 'crops/s'


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1590.0, 1400.0, 'fallback'


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled through NVML
    every 5 ms (nvidia-smi itself takes ~100 ms per query, longer than a short timed region)."""

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu, self.samples, self._stop_evt = gpu_index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        try:
            pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1e3
        except Exception:
            pw = None
        return mhz, r, pw

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.gpu)],
                                         capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(',')]
                    bits = 0
                    for bit, v in zip((0x8, 0x40, 0x20, 0x4), f[3:7]):
                        if v.lower().startswith('active'):
                            bits |= bit
                    self.max_mhz = float(f[1])
                    self.samples.append((float(f[0]), bits, float(f[2])))
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        names = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}
        reasons = set()
        for _, r, _ in self.samples:
            for bit, name in names.items():
                if r & bit:
                    reasons.add(name)
        sm = [s[0] for s in self.samples]
        pw = [s[2] for s in self.samples if s[2] is not None]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(getattr(self, 'max_mhz', 0)) or None,
                'reasons': sorted(reasons), 'samples': len(self.samples), 'power_w_max': max(pw) if pw else None}


def cpu_oracle_rate(spec, weights, dataset, n_sample, threads, repeats=1):
    """crops/s of the CPU restatement of the reference graph (torch-CPU fp32) -- NOT TF 1.13."""
    import torch
    from oracle.metro_oracle import OracleNet      # the one place bench.py may execute oracle/
    torch.set_num_threads(threads)
    net = OracleNet(spec, weights, export_permutation(dataset), 'fp32')
    img = synth_images(n_sample, seed=7)
    net(img[:1])                                   # warm-up (oneDNN primitive creation)
    t0 = time.perf_counter()
    for _ in range(repeats):
        net(img)
    dt = (time.perf_counter() - t0) / repeats
    return n_sample / dt, dt


def bind_to_gpu_numa_node(local_rank):
    """Pinned staging buffers should live on the socket the GPU hangs off (round-1 finding: eight ranks pinning on one
    node share that node's memory bandwidth).  Binds this process to the CPUs of the GPU's NUMA node before any pinned
    allocation; returns a short description for the JSON line (None if the platform exposes no topology)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), 'pci_domain_id', 0)
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f'/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node'
        node = int(open(path).read().strip())
        if node < 0:
            return 'numa_node=-1 (single node)'
        cpus = open(f'/sys/devices/system/node/node{node}/cpulist').read().strip()
        ids = set()
        for part in cpus.split(','):
            a, _, b = part.partition('-')
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return f'numa_node={node} cpus={cpus}'
    except Exception as e:           # no sysfs / no permission: leave the affinity alone
        return None


def agree_on_time_and_steps(ms_local, steps, world, dist=None, device=None):
    """Max over ranks of the per-step time and ONE step count for the clock-sampling extension of the timed loop
    (>= 1 s of steps).  Every step is a collective, so the count must be identical on every rank: it is derived from the
    all-reduced time and then broadcast from rank 0 (a rank-local count hung an 8-GPU run of this bench in round 2)."""
    import torch
    ms = ms_local
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ext = max(steps, min(2000, int(1.0 / max(ms * 1e-3, 1e-4))))
    if world > 1:
        t = torch.tensor([ext], device=device, dtype=torch.int64)
        dist.broadcast(t, 0)
        ext = int(t.item())
    return ms, ext


def make_config(cfg_name, arch, stride, j, per_gpu, world):
    """The workload description both arms print (the driver compares them)."""
    return {'workload': f'config {cfg_name}: {arch} stride_{stride} {j} joints, batch {per_gpu}/GPU',
            'global_batch': per_gpu * world, 'parallelism': f'dp{world}'}


REF_SAMPLE = 64     # crops per step of the CPU arms: a bounded sample of the workload (same as cpu_baseline)


def run_reference(args, cfg_name, arch, stride, dataset, batch):
    """--impl reference: the reference's CPU path.  TensorFlow 1.13.1 cannot be installed here
    (BASELINE.md section 4), so this times the oracle port of the same graph on all host cores, each step a
    bounded sample (REF_SAMPLE crops) of the configuration's batch."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', '1'))
    j = model_joint_info(dataset).n_joints
    spec = NetSpec(arch, stride, j)
    w = synth_weights(spec, 0)
    cores = os.cpu_count() or 1
    sample = max(1, min(batch, REF_SAMPLE if spec.flops_per_crop < 1e11 else 8))
    import torch
    from oracle.metro_oracle import OracleNet
    torch.set_num_threads(cores)
    net = OracleNet(spec, w, export_permutation(dataset), 'fp32')
    img = synth_images(sample, seed=7)
    for _ in range(max(1, min(args.warmup, 2))):
        net(img[:1])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        net(img)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': make_config(cfg_name, arch, stride, j, batch, world),
        'sample_crops_per_step': sample,
        'note': f'CPU restatement of the reference graph (torch-CPU fp32 oracle port), not TF 1.13; each step is a bounded '
                f'sample of {sample} crops of the {batch}-crop batch (rate in crops/s is per crop, so comparable)',
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{sample} crops per step x {args.steps} steps of config {cfg_name}'},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='B', choices=list(CONFIGS))
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the config\'s)')
    ap.add_argument('--head-dtype', default='f32', choices=['f32', 'f16'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--layers', action='store_true', help='also print per-layer timings to stderr')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    arch, stride, dataset, cfg_batch, cfg_gpus = CONFIGS[args.config]
    per_gpu = args.batch or (cfg_batch // max(cfg_gpus, 1) if cfg_gpus > 1 else cfg_batch)
    if args.impl == 'reference':
        return run_reference(args, args.config, arch, stride, dataset, per_gpu)

    import torch
    import torch.distributed as dist
    from metro_pose3d_b200.dist import ShardedPoseEstimator
    from metro_pose3d_b200.inference import MetroModel, SoftArgmax

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a B200: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    hbm_peak, tf_burst, tf_sust, peak_src = load_peaks()

    j = model_joint_info(dataset).n_joints
    spec = NetSpec(arch, stride, j)
    weights = synth_weights(spec, 0)
    model = MetroModel(arch, stride, dataset, weights=weights, max_batch=per_gpu, device=local_rank,
                       head_dtype=args.head_dtype)
    est = ShardedPoseEstimator(model.infer, model.n_joints_out)
    n = per_gpu
    # synthetic crops, uniform [0,1), generated ON the device, seed 1000 + rank.  Two input
    # sets are rotated; one set (n * 786 KB = 201 MB at n=256) already exceeds the 126 MB L2.
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    inputs = [torch.rand((n, 256, 256, 3), generator=g, device=dev, dtype=torch.float32) for _ in range(2)]
    outs = [torch.empty((n, model.n_joints_out, 3), dtype=torch.float32, device=dev) for _ in range(2)]

    def step(i):
        # the all-gather of step i runs on the estimator's side stream underneath the convolutions of step i + 1
        # (two result buffers); est.wait() joins it
        local = model.infer(inputs[i & 1], out=outs[i & 1])
        return est.gather_async(local, n * world)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        res = step(i)
    est.wait()                                       # the last gather joins the launch stream before the end event
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1) / args.steps
    ms_local = ms                                   # this rank's device time per step
    # the timed region above is exactly `steps` steps (~0.1 s: a handful of NVML samples); the same loop is repeated
    # for >= 1 s purely to sample clocks / throttle reasons under sustained load -- its time is reported as an extra
    ms, ext_steps = agree_on_time_and_steps(ms_local, args.steps, world, dist, dev)
    value = n * world / (ms * 1e-3)
    sampler2 = ClockSampler(local_rank)
    if rank == 0:
        sampler2.start()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for i in range(ext_steps):
        step(i)
    est.wait()
    ev3.record()
    torch.cuda.synchronize()
    clocks_ext = sampler2.stop() if rank == 0 else None
    ms_ext = ev2.elapsed_time(ev3) / ext_steps
    if world > 1:
        dist.barrier()

    # ---- end to end through the host-buffer C-ABI call: pinned host float32 in, host poses out ----
    host_in = torch.rand((n, 256, 256, 3), dtype=torch.float32).pin_memory()
    host_out = torch.empty((n, model.n_joints_out, 3), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        model.infer_host(host_in, host_out)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.infer_host(host_in, host_out)
        if world > 1:
            est.gather(torch.from_numpy(host_out.numpy()).to(dev), n * world)
            torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {'value': n * world / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(host_in.numel() * 4),
           'd2h_bytes_per_step': int(host_out.numel() * 4), 'ms_per_step': e2e_ms}
    # the same call fed uint8 crops (metro_infer_host_u8): extra information, not the headline
    host_u8 = torch.randint(0, 256, (n, 256, 256, 3), dtype=torch.uint8).pin_memory()
    for _ in range(2):
        model.infer_host(host_u8, host_out)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.infer_host(host_u8, host_out)
        if world > 1:
            est.gather(torch.from_numpy(host_out.numpy()).to(dev), n * world)
            torch.cuda.synchronize()
    u8_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    if world > 1:
        t = torch.tensor([u8_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u8_ms = float(t.item())
    e2e_u8 = {'value': n * world / (u8_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(host_u8.numel()),
              'd2h_bytes_per_step': int(host_out.numel() * 4), 'ms_per_step': u8_ms}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (conv_gemm_kernel: every tcgen05 convolution launch) ----
    # per-launch device times come from CUDA events recorded between launches on the launch stream
    model.profile(inputs[0])
    # (events between launches switch off the programmatic overlap of consecutive kernels and expose host jitter:
    # the per-launch figure is the minimum over the repetitions, and their sum is checked against the whole step)
    acc = {}
    reps = 5
    for _ in range(reps):
        for name, t_ms in model.profile(inputs[1]):
            acc[name] = min(acc.get(name, float('inf')), t_ms)
    non_gemm = ('img_pack', 'conv1+pool1', 'softargmax')
    conv_ms_launches = sum(v for k, v in acc.items() if k not in non_gemm)
    # the same launches inside the real step: whole-step device time (the timed region above) minus the other kernels
    conv_ms_step = ms_local - sum(acc[k] for k in non_gemm if k in acc)
    # ONE fixed method: the sum of the per-launch device times, each launch timed INSIDE the real pipelined step by
    # device-side stamps (earliest CTA start after its dependency wait -> latest CTA end, %globaltimer; metro_profile),
    # minimum over the repetitions.  (CUDA events between launches, the round-1 method, serialise the launches and add a
    # launch latency to each: their sum exceeded the whole step.)  The step-derived figures are named extras.
    conv_ms = conv_ms_launches
    conv_timing = 'sum of per-launch device time stamps inside the pipelined step (min over %d repetitions)' % reps
    gemm_convs = [c for c in spec.convs if c.name != 'conv1']
    # algorithmic FLOPs: 2*Ho*Wo*Cout*Cin*k^2 per conv per crop; the root conv1 has its own fused kernel
    gemm_flops = sum(c.flops for c in gemm_convs) * n
    achieved_tf = gemm_flops / (conv_ms * 1e-3) / 1e12
    # DRAM traffic of the same launches from the committed ncu capture (profiles/traffic_<config>.json), per step
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', f'traffic_{args.config}.json')
    if os.path.exists(tpath) and n == cfg_batch // max(cfg_gpus, 1):
        try:
            traffic = json.load(open(tpath)).get('conv_gemm_dram_bytes_per_step')
        except Exception:
            traffic = None
    roofline = {'kernel': 'conv_gemm_kernel + conv_chain_kernel (tcgen05 implicit GEMM; all %d convolution launches of a step)' % len([k for k in acc if k not in non_gemm]),
                'bound': 'tensor', 'achieved': achieved_tf, 'peak': tf_sust, 'unit': 'TFLOP/s',
                'frac': achieved_tf / tf_sust, 'traffic': traffic, 'peak_source': f'{peak_src} bf16 sustained',
                'algorithmic_flops_per_step': gemm_flops,
                'ms_per_step': conv_ms, 'timing': conv_timing, 'ms_sum_of_launches': conv_ms_launches,
                'extra_ms_step_minus_others': conv_ms_step,
                'extra_frac_step_minus_others': gemm_flops / (conv_ms_step * 1e-3) / 1e12 / tf_sust,
                'extra_frac_conv_flops_over_whole_step': gemm_flops / (ms_local * 1e-3) / 1e12 / tf_sust,
                'other_ms': {k: acc[k] for k in non_gemm if k in acc}}
    if args.layers:
        flops = {c.name: c.flops for c in spec.convs}
        for k, v in acc.items():
            print(f'{k:28s} {v*1e3:9.1f} us', file=sys.stderr)

    # ---- stand-alone soft-argmax HBM roofline (second half of the BASELINE metric) ----
    side = spec.feat_side
    perm = export_permutation(dataset)
    sam = SoftArgmax(side, j, stride, perm, head_dtype=args.head_dtype)
    isz = 4 if args.head_dtype == 'f32' else 2
    heads = []
    n_rot = max(2, int(np.ceil(300e6 / (n * side * side * 8 * j * isz))))   # rotate > 2x L2 worth of inputs
    hx = torch.from_numpy(synth_head(min(n, 32), side, j, seed=0)).to(dev)
    reps_h = int(np.ceil(n / hx.shape[0]))
    for r in range(n_rot):
        t = hx.repeat(reps_h, 1, 1, 1)[:n].contiguous().roll(r, 0)
        heads.append(t.half() if args.head_dtype == 'f16' else t)
    pout = torch.empty((n, len(perm), 3), dtype=torch.float32, device=dev)
    for r in range(n_rot):
        sam(heads[r], pout)
    torch.cuda.synchronize()
    # device time per launch: the launches are replayed from a CUDA graph, so the figure is not bounded by the
    # ~10 us of Python/ctypes per call (the kernel itself runs for about as long)
    iters = 4 * n_rot
    side_stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side_stream):
        with torch.cuda.graph(graph, stream=side_stream):
            for i in range(iters):
                sam(heads[i % n_rot], pout)
        graph.replay()
        torch.cuda.synchronize()
        reps_g = 5
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(side_stream)
        for _ in range(reps_g):
            graph.replay()
        s1.record(side_stream)
    torch.cuda.synchronize()
    sam_us = s0.elapsed_time(s1) / (iters * reps_g) * 1e3
    sam_bytes = spec.softargmax_bytes_per_crop(len(perm), isz) * n
    sam_gbs = sam_bytes / (sam_us * 1e-6) / 1e9
    sam_traffic = None
    if os.path.exists(tpath) and n == cfg_batch // max(cfg_gpus, 1):
        try:
            sam_traffic = json.load(open(tpath)).get('softargmax_dram_bytes_per_launch')
        except Exception:
            sam_traffic = None
    roofline_sam = {'kernel': 'softargmax_kernel', 'bound': 'hbm', 'achieved': sam_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': sam_gbs / hbm_peak, 'traffic': sam_traffic, 'us_per_launch': sam_us, 'bytes_per_launch': sam_bytes,
                    'note': f'{n_rot} rotating inputs ({n_rot * sam_bytes / 1e6:.0f} MB > L2), back-to-back launches replayed from a CUDA graph'}

    # ---- CPU baseline beside it (bounded sample; reported, not the target) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample, passes = REF_SAMPLE, 3             # ~10-20 s of CPU work on the box's host cores
        rate, dt = cpu_oracle_rate(spec, weights, dataset, sample, cores, repeats=passes)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': f'{sample} crops x {passes} passes of config {args.config} ({dt:.2f} s/pass), torch-CPU fp32 oracle port (not TF 1.13)'}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f16', 'data': 'synthetic',
        'config': make_config(args.config, arch, stride, j, n, world),
        'details': {'l2': 'inputs larger than L2 (2 rotating batches)',
                    'arithmetic': 'f16 operands (the reference default, src/options.py:73), f32 accumulate, f32 head and decode',
                    'head_dtype': args.head_dtype, 'gflop_per_crop': spec.flops_per_crop / 1e9,
                    'tensor_frac_whole_step': spec.flops_per_crop * n / (ms_local * 1e-3) / 1e12 / tf_sust,
                    'host_numa_binding': numa,
                    'extended_region': {'steps': ext_steps, 'ms_per_step': ms_ext, 'clocks': clocks_ext}},
        'e2e': e2e, 'e2e_u8': e2e_u8, 'gpu_launches': model.launch_count(n) * args.steps,
        'roofline': roofline, 'roofline_softargmax': roofline_sam, 'cpu_baseline': cpu, 'clocks': clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
