#!/usr/bin/env python
"""Benchmark of the AttentionShift hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path   (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (CPU oracle port)

One "step" = one pass of the hot path over one batch of synthetic input:
    VisionTransformerDet.forward (ViT-B/16, 12 blocks, head-mean attention maps of the last 7 layers)
    -> AttnShiftRoIHead.seed_pseudo_gt (roll-out slab, CAM boxes, refined instance maps, mask points,
       mean-shift part discovery, pseudo masks)
Workload = BASELINE.json configs[1]: bs8 1024x1024 ViT-B/16, 5 attention-shift iterations, 16 seeds / instance, per GPU
(weak scaling: every rank runs its own batch of 8; no data-path collective -- SURVEY.md 8e).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs.  cfg2 is the one the metric is quoted on (default); cfg1 is the reference's own CPU-runnable case,
# cfg3 the roofline-capture case, cfg4 one rank's share of the 8-GPU DDP case, cfg5 the ViT-L / COCO-shape case (2 per GPU).
CONFIGS = {
    'cfg1': dict(name='cfg1: 1x224x224 ViT-B/16, 2 attn-shift iters, 4 seeds/instance', batch=1, img=(224, 224), embed=768, depth=12,
                 heads=12, n_obj=2, iters=2, seeds=4),
    'cfg2': dict(name='cfg2: bs8 1024x1024 ViT-B/16, 5 attn-shift iters, 16 seeds/instance', batch=8, img=(1024, 1024), embed=768,
                 depth=12, heads=12, n_obj=3, iters=5, seeds=16),
    'cfg3': dict(name='cfg3: bs32 1024x1024 ViT-B/16, 10 attn-shift iters, 32 seeds/instance', batch=32, img=(1024, 1024), embed=768,
                 depth=12, heads=12, n_obj=3, iters=10, seeds=32),
    'cfg4': dict(name='cfg4: bs64 over 8 GPUs = bs8 per GPU, 1024x1024 ViT-B/16, 10 attn-shift iters, 16 seeds/instance', batch=8,
                 img=(1024, 1024), embed=768, depth=12, heads=12, n_obj=3, iters=10, seeds=16),
    'cfg5': dict(name='cfg5: bs16 over 8 GPUs = bs2 per GPU, 1344x800 (COCO shape, padded) ViT-L/16, 10 attn-shift iters, 64 seeds/instance',
                 batch=2, img=(800, 1344), embed=1024, depth=24, heads=16, n_obj=3, iters=10, seeds=64),
}
for _c in CONFIGS.values():
    _c.update(cam_layer=7, n_point_tokens=100)
WORKLOAD = CONFIGS['cfg2']


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg2', choices=sorted(CONFIGS))
    ap.add_argument('--mode', default='forward', choices=['forward', 'train'],
                    help='forward (default, the BASELINE metric): backbone forward + seed_pseudo_gt; train: + backward of the backbone, '
                         'DDP gradient all-reduce (NCCL) and the optimizer step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-drop-path', action='store_true', help='--mode train: stochastic depth off (the shipped config trains with 0.05)')
    ap.add_argument('--no-reference-config', action='store_true', help='skip the untouched-reference-config leg (cfg2, N=1)')
    ap.add_argument('--small', action='store_true', help='tiny config for a functional check (not a valid bench number)')
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed regions (profiling recipe): the same NVML counters nvidia-smi prints,
    read in-process.  (Spawning nvidia-smi every 200 ms forks a process that maps tens of GB of CUDA memory; the forks
    stalled the launching thread for tens of milliseconds and showed up as 30-70 ms steps.)  Falls back to the
    nvidia-smi query of the recipe when pynvml is missing."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index, period=0.5):
        super().__init__(daemon=True)
        self.index = index
        self.period = period
        self.rows = []           # (sm_mhz, sm_max_mhz, [reasons])
        self.stop_flag = False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and all(v.strip().isdigit() for v in vis.split(',')) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, 'nvmlDeviceGetCurrentClocksEventReasons') \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        names = []
        for name, bit in (('hw_slowdown', n.nvmlClocksThrottleReasonHwSlowdown), ('hw_thermal_slowdown', n.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ('sw_thermal_slowdown', n.nvmlClocksThrottleReasonSwThermalSlowdown), ('sw_power_cap', n.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                names.append(name)
        self.rows.append((float(sm), float(mx), names))

    def _sample_smi(self):
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            r = [x.strip() for x in out.split(',')]
            names = [nm for nm, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[3:7])
                     if v.lower().startswith('active')]
            self.rows.append((float(r[0]), float(r[1]), names))

    def run(self):
        while not self.stop_flag:
            try:
                self._sample_nvml() if self.nvml is not None else self._sample_smi()
            except Exception:
                pass
            time.sleep(self.period if self.nvml is not None else 1.0)

    def summary(self):
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({nm for r in self.rows for nm in r[2]})
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.rows[0][1], reasons=reasons, samples=len(self.rows),
                    source='nvml' if self.nvml is not None else 'nvidia-smi')


def usable_cpus():
    """Host cores this process may actually use: the affinity mask and the cgroup CPU quota, not the machine's core count
    (128 torch threads on a container throttled to a few cores made the CPU arm 10x slower than 16 threads on 16 cores)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for path in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
        try:
            txt = open(path).read().split()
            if path.endswith('cpu.max'):
                quota, period = txt[0], float(txt[1])
            else:
                quota, period = txt[0], float(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
            if quota not in ('max', '-1'):
                n = min(n, max(1, int(float(quota) / period + 0.5)))
            break
        except (OSError, ValueError, IndexError):
            continue
    return max(1, n)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sus=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src='fallback')


# ----------------------------------------------------------------------------------------------------------------- ours
def make_inputs(cfg, rank, pin=True):
    """Synthetic batch of the workload: randn images (already normalised), GT points, and the two selections that stand in for
    the learned / host-side stages: pos_inds for the Hungarian match, gt_index for the MIL layer choice.  The SAME
    function feeds the CUDA arm and the CPU arm (which runs image 0 of rank 0's batch)."""
    g = torch.Generator().manual_seed(1234 + rank)
    B, (H, W) = cfg['batch'], cfg['img']
    img = torch.randn(B, 3, H, W, generator=g)
    if pin:
        img = img.pin_memory()
    n = cfg['n_obj']
    span = torch.tensor([W - 0.2 * W, H - 0.2 * H])
    gt_points = [(torch.rand(n, 2, generator=g) * span + 0.1 * torch.tensor([W, H])).floor() for _ in range(B)]
    pos_inds = [torch.arange(n) for _ in range(B)]                       # stands in for the Hungarian match
    gt_index = [torch.randint(0, cfg['cam_layer'], (n,), generator=g) for _ in range(B)]    # stands in for the MIL choice
    labels = [torch.randint(0, 20, (n,), generator=g) for _ in range(B)]
    return img, gt_points, pos_inds, gt_index, labels


def build_models(cfg, dev, reference_config=False):
    """reference_config=False: the fast drop-in -- head-mean maps only for the 7 layers seed_pseudo_gt reads and only in the form the
    roll-out consumes, no FPN (outside SURVEY 8a), forward replayed as one CUDA graph.  reference_config=True: the backbone exactly
    as configs/mae builds it (all 12 full fp32 maps, FPN on, eager launches)."""
    from attentionshift_b200.registry import build_backbone, build_head
    from attentionshift_b200.synthetic import vit_state_dict
    H, W = cfg['img']
    kw = dict(with_fpn=True) if reference_config else dict(with_fpn=False, attn_layers=cfg['cam_layer'], attn_format='rollout',
                                                            cuda_graph=cfg.get('cuda_graph', True))
    bb = build_backbone(dict(type='VisionTransformerDet', img_size=H if H == W else 224, patch_size=16, embed_dim=cfg['embed'],
                             depth=cfg['depth'], num_heads=cfg['heads'], mlp_ratio=4, qkv_bias=True, last_feat=True,
                             return_attention=True, point_tokens_num=cfg['n_point_tokens'], out_indices=[3, 5, 7, 11],
                             drop_path_rate=cfg.get('drop_path_rate', 0.0), **kw))
    sd = vit_state_dict(cfg['embed'], cfg['depth'], cfg['heads'], H if H == W else 224, n_point_tokens=cfg['n_point_tokens'], seed=0)
    bb.load_state_dict(sd, strict=False)
    bb = bb.to(dev).eval()
    head = build_head(dict(type='AttnShiftRoIHead', bbox_head=dict(cam_layer=cfg['cam_layer'], seed_thr=0.2, seed_multiple=0.5),
                           mean_shift_times_local=cfg['iters'], n_seeds=cfg['seeds'], num_semantic_points=3))
    return bb, head


def one_step(bb, head, img_dev, inputs, return_mask):
    _, gt_points, pos_inds, gt_index, labels = inputs
    out = bb(img_dev)
    hp, wp = img_dev.shape[-2] // 16, img_dev.shape[-1] // 16
    vit_feat = out['last_feat'][:, 1:]                                   # token-major [B,N,C] view (DET:77 without the transpose)
    res = head.seed_pseudo_gt(None, None, None, None, None, vit_feat=vit_feat.unflatten(1, (hp, wp)).permute(0, 3, 1, 2),
                              point_cls=out['outputs_class'], point_reg=out['outputs_coord'], attns=out['attns'],
                              gt_points=gt_points, gt_points_labels=labels, return_mask=return_mask, pos_mask_thr=0.6,
                              neg_mask_thr=0.1, num_mask_point_gt=10, corr_size=21, obj_tau=0.85, pos_inds=pos_inds,
                              gt_index=gt_index)
    return res


OURS = ('linear_tcgen05', 'mhsa_fwd', 'attn_headmean', 'layernorm_f16', 'im2col16', 'assemble_tokens', 'rollout_', 'cam_', 'ccl_',
        'ms_', 'mean_shift', 'split_tokens', 'norm_', 'seed_proto', 'refine_', 'weighted_sum', 'cos_warp', 'fuse_', 'crop_', 'erode_down',
        'filter_score', 'merge_protos', 'part_', 'fill_u32', 'minmax_decode', 'ext_init')


def count_launches(fn):
    """Number of kernels launched by this repo's .so in one step (CUPTI via torch.profiler)."""
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        n_ours = n_all = 0
        for e in prof.events():
            if 'cuda' in str(e.device_type).lower() and e.name and not e.name.startswith('Memcpy') and not e.name.startswith('Memset'):
                n_all += 1
                if any(k in e.name for k in OURS):
                    n_ours += 1
        return n_ours, n_all
    except Exception:
        return None, None


def pregrow_allocator(dev, mib=96):
    """Give torch's caching allocator a cushion of small-pool segments before the timed regions.  The head allocates a few hundred
    small tensors per step; whenever their live set exceeds what the small pool has cached, the allocator grows it by one 2 MiB
    segment = a cudaMalloc in the middle of a step, which showed up as ONE 40-110 ms step in some runs of ten
    (`e2e.allocator_growth_mib` reports any growth that still happens inside the timed e2e region)."""
    hold = [torch.empty(512 * 1024, device=dev, dtype=torch.uint8) for _ in range(2 * mib)]
    del hold


def time_steps(fn, steps, barrier, per_step=None):
    """K steps bracketed by barrier + synchronize on both sides, CUDA events on the launching stream.  ``per_step``: list that
    receives the individual step times (an event after every step: no synchronisation, the timed region is unchanged)."""
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    gc.collect()
    gc.disable()            # like timeit: a generation-2 collection of the interpreter (tens of ms with a model's object graph alive)
    barrier()               # inside a 25 ms step is the host's noise, not the path's -- it showed up as one 40-60 ms step in ten
    evs[0].record()
    for i in range(steps):
        fn()
        evs[i + 1].record()
    barrier()
    gc.enable()
    if per_step is not None:
        per_step.extend(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
    return evs[0].elapsed_time(evs[steps]) / steps


_REAL_STDOUT = None


def protect_stdout():
    """stdout carries exactly ONE line, the JSON record.  Libraries write there too (NCCL prints its version / NCCL_DEBUG=INFO log
    on fd 1): point fd 1 at stderr for the duration of the run -- the log stays visible for rank / transport checks -- and keep
    the real stdout for ``emit``."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def run_ours(args):
    from attentionshift_b200 import parallel
    protect_stdout()
    rank, world, local = parallel.env_rank_world()
    # the host side of a rank is one launching thread: keep torch's CPU pool from oversubscribing the box when 8 ranks share it
    torch.set_num_threads(max(1, min(4, usable_cpus() // max(world, 1))))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    parallel.init('nccl', dev)
    cfg = dict(CONFIGS[args.config])
    if args.small:
        cfg.update(batch=2, img=(224, 224), depth=7)
    from attentionshift_b200 import ops
    bb, head = build_models(cfg, dev)
    inputs = make_inputs(cfg, rank)
    img_host = inputs[0]
    img_dev = img_host.to(dev)

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize()

    # the clock sampler starts BEFORE the warm-up: the first NVML queries of a process are slow and take a driver lock that
    # stalls kernel launches (seen as a 100+ ms hiccup in whichever loop ran first); its samples are reset when timing starts
    sampler = ClockSampler(local) if rank == 0 else None     # one per job: NVML calls serialise on a driver lock shared by all ranks
    if os.environ.get('AS_BENCH_NO_SAMPLER'):
        sampler = None                                       # diagnosis only: a line without `clocks` is not a valid bench line
    if sampler is not None:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        one_step(bb, head, img_dev, inputs, False)
    torch.cuda.synchronize()
    pregrow_allocator(dev)

    # ---- device-resident timing (value): library timing slots OFF, nothing but the step's own work on the stream
    if sampler is not None:
        sampler.rows = []
    step_ms = []
    ms_dev = time_steps(lambda: one_step(bb, head, img_dev, inputs, False), args.steps, barrier, step_ms)

    # ---- end-to-end timing (e2e): pinned host image -> device every step, masks back to the host every step.
    # The copy of step i+1 runs on a side stream while step i computes (double buffer); every copy is inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(img_dev), torch.empty_like(img_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i % 2])
            bufs[i % 2].copy_(img_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_loop(n, marks=None):
        d2h = 0
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            res = one_step(bb, head, bufs[i % 2], inputs, True)
            freed[i % 2].record()
            d2h = sum(m.nbytes for m in res['pseudo_gt_masks'])
            if marks is not None:
                marks[i + 1].record()
        return d2h

    for f in freed:
        f.record()
    e2e_loop(warm)                         # untimed: first use of the pinned mask buffers / copy stream (cudaHostAlloc is slow)
    gc.collect()
    gc.disable()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    mem0 = torch.cuda.memory_reserved()
    e2.record()
    marks[0].record()
    d2h = e2e_loop(args.steps, marks)
    e3.record()
    barrier()
    gc.enable()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    e2e_step_ms = [round(marks[i].elapsed_time(marks[i + 1]), 2) for i in range(args.steps)]
    mem_growth_mib = (torch.cuda.memory_reserved() - mem0) / 2 ** 20
    if sampler is not None:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    ms_dev, ms_e2e = parallel.max_over_ranks([ms_dev, ms_e2e], device=dev)     # the slowest rank defines the step

    # ---- per-entry-point device times: a SEPARATE, untimed pass with the library's timing slots on (they are event-record
    # nodes inside a re-captured backbone graph and ~100 extra event records per step: kept out of the timed regions above)
    fam = {}
    n_ours = n_all = None
    if rank == 0:
        ops.TIMERS.enable()
        if hasattr(bb, '_graphs'):
            bb._graphs.clear()
        for _ in range(3):
            ops.TIMERS.begin_step()
            one_step(bb, head, img_dev, inputs, False)
        fam = ops.TIMERS.summary()
        ops.TIMERS.disable()
        if hasattr(bb, '_graphs'):
            bb._graphs.clear()
        n_ours, n_all = count_launches(lambda: one_step(bb, head, img_dev, inputs, False))

    # ---- the reference's configuration untouched (all 12 full fp32 maps, FPN on, eager launches): cfg2, one GPU
    ref_cfg = None
    if rank == 0 and world == 1 and args.config == 'cfg2' and not args.no_reference_config and not args.small:
        del bufs
        bb._graphs.clear()
        torch.cuda.empty_cache()
        bb2, head2 = build_models(cfg, dev, reference_config=True)
        for _ in range(2):
            one_step(bb2, head2, img_dev, inputs, False)
        ms2 = time_steps(lambda: one_step(bb2, head2, img_dev, inputs, False), max(2, min(args.steps, 4)), lambda: torch.cuda.synchronize())
        ref_cfg = dict(value=round(cfg['batch'] / (ms2 * 1e-3), 2), unit='images/s', ms_per_step=round(ms2, 3),
                       what='VisionTransformerDet exactly as configs/mae/attnshift_voc12aug.py builds it: all 12 head-mean maps as full fp32 '
                            '[B,T,T] tensors (attn_layers=None, attn_format=full), with_fpn=True, eager launches (cuda_graph=False)')
        del bb2, head2

    if rank == 0:
        pk = peaks()
        H, W = cfg['img']
        N = (H // 16) * (W // 16)
        B, T, C = cfg['batch'], 1 + N + cfg['n_point_tokens'], cfg['embed']
        att = fam.get('as_mhsa_fwd', {})
        flops_attn = 4.0 * T * T * C * B                      # SDPA part of F_attn, per launch (one layer, whole batch)
        roof = None
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'ncu_mhsa_traffic.json')      # dram bytes / launch from the committed ncu capture
        if os.path.exists(tpath) and not args.small and args.config == 'cfg2':
            try:
                traffic = json.load(open(tpath)).get('dram_bytes_per_launch')
            except Exception:
                traffic = None
        if att.get('n'):
            ach = flops_attn / (att['ms'] / att['n'] * 1e-3) / 1e12
            roof = dict(kernel='mhsa_fwd2_kernel (tcgen05 flash attention, one launch = one layer x batch)', bound='tensor',
                        achieved=round(ach, 1), peak=pk['tf_sus'], unit='TFLOP/s', frac=round(ach / pk['tf_sus'], 4),
                        traffic=traffic, peak_source=pk['src'] + ' sustained bf16 GEMM', share_of_step=round(att['ms'] / ms_dev, 3))
        ms_name = next((k for k in ('as_mean_shift_v2', 'as_mean_shift_fused', 'as_mean_shift_tc', 'as_mean_shift') if fam.get(k, {}).get('n')),
                       'as_mean_shift')
        msf = fam.get(ms_name, {})
        K = cfg['n_obj'] * cfg['seeds']
        b_alg = ((cfg['iters'] + 1) * N * C * 4 + K * N * 4 + 2 * K * C * 4) * B      # SURVEY 8d B_alg per image x images
        roof2 = None
        if msf.get('n'):
            ach2 = b_alg / (msf['ms'] / msf['n'] * 1e-3) / 1e9
            what = {'as_mean_shift_v2': 'one persistent cooperative kernel, two CTAs per SM, + the token split kernel',
                    'as_mean_shift_fused': 'persistent cooperative kernel of round 1 with the round-2 changes (8 x 8 patch-block units, unit skipping) + the token split kernel',
                    'as_mean_shift_tc': '~50 launches', 'as_mean_shift': 'fp32 CUDA-core kernels'}[ms_name]
            traffic2 = None
            t2path = os.path.join(ROOT, 'profiles', {'as_mean_shift_v2': 'ncu_msv2_traffic.json',
                                                     'as_mean_shift_fused': 'ncu_msfused_traffic.json'}.get(ms_name, 'none'))
            if os.path.exists(t2path) and not args.small and args.config == 'cfg2':      # dram bytes / launch from the committed ncu capture
                try:
                    traffic2 = json.load(open(t2path)).get('dram_bytes_per_launch')
                except Exception:
                    traffic2 = None
            roof2 = dict(kernel='%s (whole on-device attention-shift loop, all images of the batch; %s)' % (ms_name, what), bound='hbm',
                         achieved=round(ach2, 1), peak=pk['hbm'], unit='GB/s', frac=round(ach2 / pk['hbm'], 4), traffic=traffic2,
                         peak_source=pk['src'], ms_per_call=round(msf['ms'] / msf['n'], 4), algorithmic_bytes=int(b_alg))
        line = dict(metric='images/sec at 1024^2 bs8 ViT-B attn-shift' if args.config == 'cfg2' else 'images/sec, ' + cfg['name'],
                    value=round(world * B / (ms_dev * 1e-3), 2), unit='images/s',
                    n_gpus=world, steps=args.steps, warmup=warm, ms_per_step=round(ms_dev, 3), step_ms=[round(v, 2) for v in step_ms],
                    higher_is_better=True,
                    scaling='weak', vs_baseline=None, dtype='f16 operands / f32 accumulate (ViT GEMMs + attention), f32 (attention shift)',
                    data='synthetic (random-init ViT weights, randn images, random GT points)',
                    config=dict(workload=cfg['name'], per_gpu_batch=B, mode='forward-only, no collective (every rank runs its own batch)',
                                fpn=False, l2='inputs larger than L2 (per-step working set >> 126 MB)',
                                backbone='one CUDA graph per forward (cuda_graph=True); head-mean maps of the 7 layers seed_pseudo_gt reads, '
                                         'produced as roll-out operands (attn_layers=7, attn_format=rollout)',
                                kernel_times='library timing slots, separate untimed pass after the timed regions', small=bool(args.small)),
                    e2e=dict(value=round(world * B / (ms_e2e * 1e-3), 2), unit='images/s', ms_per_step=round(ms_e2e, 3),
                             h2d_bytes_per_step=int(img_host.nbytes), d2h_bytes_per_step=int(d2h), step_ms=e2e_step_ms,
                             allocator_growth_mib=round(mem_growth_mib, 1)),
                    gpu_launches=n_ours, all_launches=n_all, clocks=sampler.summary() if sampler is not None else None, roofline=roof, roofline_attnshift=roof2,
                    reference_config=ref_cfg,
                    kernel_ms_per_step={k: round(v['ms'], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]['ms'])})
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline(cfg, steps=1)
        emit(line)
    if world > 1:
        parallel.barrier()
        torch.distributed.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------- CPU arm
def cpu_one_image(cfg, which=0):
    """The reference algorithm (CPU oracle port, same torch ops as the reference) on ONE image of the workload -- image ``which`` of
    the very batch the CUDA arm runs (same generator), the same dataflow (backbone outputs -> roll-out -> CAM boxes -> attention
    shift), the same selections and the same seed count."""
    from attentionshift_b200.synthetic import vit_state_dict
    from oracle import attnshift as O
    from oracle import vit as V
    H, W = cfg['img']
    hp, wp = H // 16, W // 16
    n_pt = cfg['n_point_tokens']
    if not hasattr(cpu_one_image, 'cache') or cpu_one_image.cache[0] != cfg['name']:
        sd = vit_state_dict(cfg['embed'], cfg['depth'], cfg['heads'], H if H == W else 224, n_point_tokens=n_pt, seed=0)
        cpu_one_image.cache = (cfg['name'], sd, make_inputs(cfg, 0, pin=False))
    _, sd, (img, gt_points, pos_inds, gt_index, labels) = cpu_one_image.cache
    i = which % cfg['batch']
    n = cfg['n_obj']
    t0 = time.time()
    with torch.no_grad():
        out = V.backbone_forward(img[i:i + 1], sd, cfg['depth'], cfg['heads'], n_point_tokens=n_pt)
        t1 = time.time()
        rows = O.rollout_rows(out['attns'][-cfg['cam_layer']:], n_pt)[0]      # slab shortcut (100 rows), NOT the reference's full T x T chain (RH:1265)
        low, up = O.cams_from_rollout(rows, pos_inds[i], n_pt, hp, wp)
        boxes = torch.stack([torch.cat([O.bbox_from_cam(up[l, j].clone(), gt_points[i][j], 0.2, 0.5, (H, W))[0] for j in range(n)])
                             for l in range(cfg['cam_layer'])])
        pb = boxes[gt_index[i], torch.arange(n)]
        O.attention_shift_image(up, gt_index[i], pb, out['last_feat'][0, 1:].t().unflatten(-1, (hp, wp)).contiguous(), gt_points[i],
                                labels[i], pos_mask_thr=0.6, neg_mask_thr=0.1, num_mask_point_gt=10, corr_size=21, obj_tau=0.85,
                                mean_shift_times=cfg['iters'], n_points=cfg['seeds'])
    t2 = time.time()
    return t2 - t0, t1 - t0, t2 - t1


CPU_SAMPLE = ('one image of the workload batch per step (the same image, selections and {seeds} seeds / {iters} iterations as the CUDA arm): '
              'oracle port of the reference on torch CPU fp32, all usable host threads; roll-out as the 100-row slab (not the '
              "reference's full T x T chain, RH:1265, which would add ~0.8 s / image)")


def cpu_baseline(cfg, steps=1):
    torch.set_num_threads(usable_cpus())
    ts = [cpu_one_image(cfg, i) for i in range(steps)]
    tot = sum(t[0] for t in ts) / len(ts)
    return dict(value=round(1.0 / tot, 4), unit='images/s', cores=usable_cpus(), kind='port',
                sample=CPU_SAMPLE.format(**cfg) + f'; ViT forward {ts[0][1]:.1f}s + roll-out / attention shift {ts[0][2]:.1f}s')


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cfg = dict(CONFIGS[args.config])
    if args.small:
        cfg.update(batch=2, img=(224, 224), depth=7)
    torch.set_num_threads(usable_cpus())
    budget = 240.0
    t_all = time.time()
    done_w = 0
    for _ in range(min(args.warmup, 1)):
        cpu_one_image(cfg, 0)
        done_w += 1
    per = []
    for i in range(args.steps):
        if per and (time.time() - t_all) + per[-1] > budget:
            break
        per.append(cpu_one_image(cfg, i)[0])
    ms = 1e3 * sum(per) / len(per)
    v = round(1e3 / ms, 4)
    line = dict(impl='reference', metric='images/sec at 1024^2 bs8 ViT-B attn-shift' if args.config == 'cfg2' else 'images/sec, ' + cfg['name'],
                value=v, unit='images/s', n_gpus=args.gpus,
                steps=len(per), warmup=done_w, ms_per_step=round(ms, 1), higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic (the CUDA arm\'s own batch: same generator, same selections)',
                config=dict(workload=cfg['name'], per_gpu_batch=cfg['batch'], small=bool(args.small), mode='forward-only, no collective',
                            note='each step = ONE image of the batch (bounded sample); the reference algorithm is per-image, images/s is per host'),
                cpu_baseline=dict(value=v, unit='images/s', cores=usable_cpus(), kind='port',
                                  sample=CPU_SAMPLE.format(**cfg) + ' (the python reference cannot travel to the GPU box)'),
                e2e=dict(value=v, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------- training step
def run_train(args):
    """Training step of the hot path under DDP (reference: mmdet/apis/train.py:96-100 wraps the detector in
    MMDistributedDataParallel, mmdet/utils/optimizer.py:23-38 scales the loss and steps the optimizer):
        backbone forward with autograd (device kernels) -> seed_pseudo_gt without gradient (pseudo labels, DET:75-91)
        -> surrogate loss on the backbone outputs the detector's losses consume (last_feat, point-token heads) -> backward on the
        device kernels, gradient all-reduce overlapped by DDP's bucketing -> AdamW step.
    The detector's own losses (RPN / RoI heads) are outside the hot path; the surrogate keeps every backbone parameter in the
    graph so that the all-reduce moves the full 86 M-parameter gradient (344 MB in fp32)."""
    from attentionshift_b200 import parallel
    protect_stdout()
    rank, world, local = parallel.env_rank_world()
    torch.set_num_threads(max(1, min(4, usable_cpus() // max(world, 1))))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    parallel.init('nccl', dev)
    cfg = dict(CONFIGS[args.config])
    if args.small:
        cfg.update(batch=2, img=(224, 224), depth=7)
    cfg['cuda_graph'] = False
    cfg['drop_path_rate'] = 0.0 if args.no_drop_path else 0.05     # configs/mae/attnshift_voc12aug.py:28 (stochastic depth in training)
    bb, head = build_models(cfg, dev)
    bb.train()
    model = bb
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(bb, device_ids=[local], broadcast_buffers=False, gradient_as_bucket_view=True)
    opt = torch.optim.AdamW(bb.parameters(), lr=1e-5, weight_decay=0.05, fused=True)
    inputs = make_inputs(cfg, rank)
    img_dev = inputs[0].to(dev)
    _, gt_points, pos_inds, gt_index, labels = inputs
    hp, wp = cfg['img'][0] // 16, cfg['img'][1] // 16
    n_params = sum(p.numel() for p in bb.parameters() if p.requires_grad)
    loss_scale = 1024.0                                                 # static loss scale (apex O1 scales dynamically): fp16 GEMM operands in the backward

    def step():
        out = model(img_dev)
        with torch.no_grad():
            head.seed_pseudo_gt(None, None, None, None, None, vit_feat=out['last_feat'][:, 1:].detach().unflatten(1, (hp, wp)).permute(0, 3, 1, 2),
                                point_cls=out['outputs_class'].detach(), point_reg=out['outputs_coord'].detach(), attns=out['attns'],
                                gt_points=gt_points, gt_points_labels=labels, return_mask=False, pos_mask_thr=0.6, neg_mask_thr=0.1,
                                num_mask_point_gt=10, corr_size=21, obj_tau=0.85, pos_inds=pos_inds, gt_index=gt_index)
        loss = (out['last_feat'].float().pow(2).mean() + out['outputs_class'].float().pow(2).mean()
                + out['outputs_coord'].float().pow(2).mean()) * loss_scale
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    warm = max(args.warmup, 6)          # the caching allocator needs a few full fwd + bwd steps to stop growing (each growth is a cudaMalloc)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.rows = []
    step_ms = []
    ms = time_steps(step, args.steps, barrier, step_ms)
    loss = float(step().detach()) / loss_scale
    if sampler is not None:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    ms, = parallel.max_over_ranks([ms], device=dev)
    if rank == 0:
        B = cfg['batch']
        line = dict(metric='images/sec, training step, ' + cfg['name'], value=round(world * B / (ms * 1e-3), 2), unit='images/s', n_gpus=world,
                    steps=args.steps, warmup=warm, ms_per_step=round(ms, 3), step_ms=[round(v, 2) for v in step_ms], higher_is_better=True,
                    scaling='weak', vs_baseline=None, dtype='f16 operands / f32 accumulate (GEMMs, attention fwd + bwd), f32 (LayerNorm, GELU, AdamW, attention shift)',
                    data='synthetic (random-init ViT weights, randn images, random GT points)',
                    config=dict(workload=cfg['name'], per_gpu_batch=B,
                                mode='train: backbone fwd + seed_pseudo_gt (no grad) + surrogate loss + backbone bwd + DDP gradient all-reduce (NCCL, '
                                     '%d parameters = %.0f MB fp32 per step) + fused AdamW' % (n_params, n_params * 4 / 1e6),
                                collective='torch DDP bucketed all-reduce overlapped with the backward' if world > 1 else 'none (1 rank)',
                                drop_path_rate=cfg['drop_path_rate'], loss=round(loss, 6)),
                    clocks=sampler.summary() if sampler is not None else None)
        emit(line)
    if world > 1:
        parallel.barrier()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    elif a.mode == 'train':
        run_train(a)
    else:
        run_ours(a)
