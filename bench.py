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
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name='cfg2: bs8 1024x1024 ViT-B/16, 5 attn-shift iters, 16 seeds/instance', batch=8, img=1024, embed=768,
                depth=12, heads=12, n_obj=3, iters=5, seeds=16, cam_layer=7, n_point_tokens=100)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--small', action='store_true', help='tiny config for a functional check (not a valid bench number)')
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed regions (profiling recipe): the same NVML counters nvidia-smi prints,
    read in-process.  (Spawning nvidia-smi every 200 ms forks a process that maps tens of GB of CUDA memory; the forks
    stalled the launching thread for tens of milliseconds and showed up as 30-70 ms steps.)  Falls back to the
    nvidia-smi query of the recipe when pynvml is missing."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index, period=0.25):
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
def make_inputs(cfg, rank):
    g = torch.Generator().manual_seed(1234 + rank)
    B, S = cfg['batch'], cfg['img']
    img = torch.randn(B, 3, S, S, generator=g).pin_memory()
    n = cfg['n_obj']
    gt_points = [(torch.rand(n, 2, generator=g) * (S - 200) + 100).floor() for _ in range(B)]
    pos_inds = [torch.arange(n) for _ in range(B)]                       # stands in for the Hungarian match
    gt_index = [torch.randint(0, cfg['cam_layer'], (n,), generator=g) for _ in range(B)]    # stands in for the MIL choice
    labels = [torch.randint(0, 20, (n,), generator=g) for _ in range(B)]
    return img, gt_points, pos_inds, gt_index, labels


def build_models(cfg, dev):
    from attentionshift_b200.registry import build_backbone, build_head
    from attentionshift_b200.synthetic import vit_state_dict
    bb = build_backbone(dict(type='VisionTransformerDet', img_size=cfg['img'], patch_size=16, embed_dim=cfg['embed'],
                             depth=cfg['depth'], num_heads=cfg['heads'], mlp_ratio=4, qkv_bias=True, with_fpn=False,
                             last_feat=True, return_attention=True, point_tokens_num=cfg['n_point_tokens'],
                             attn_layers=cfg['cam_layer'], out_indices=[3, 5, 7, 11], cuda_graph=cfg.get('cuda_graph', True)))
    sd = vit_state_dict(cfg['embed'], cfg['depth'], cfg['heads'], cfg['img'], n_point_tokens=cfg['n_point_tokens'], seed=0)
    bb.load_state_dict(sd, strict=False)
    bb = bb.to(dev).eval()
    head = build_head(dict(type='AttnShiftRoIHead', bbox_head=dict(cam_layer=cfg['cam_layer'], seed_thr=0.2, seed_multiple=0.5),
                           mean_shift_times_local=cfg['iters'], n_seeds=cfg['seeds'], num_semantic_points=3))
    return bb, head


def one_step(bb, head, img_dev, inputs, return_mask):
    _, gt_points, pos_inds, gt_index, labels = inputs
    out = bb(img_dev)
    hp = img_dev.shape[-1] // 16
    vit_feat = out['last_feat'][:, 1:]                                   # token-major [B,N,C] view (DET:77 without the transpose)
    res = head.seed_pseudo_gt(out['feature'], None, None, None, None, vit_feat=vit_feat.unflatten(1, (hp, hp)).permute(0, 3, 1, 2),
                              point_cls=out['outputs_class'], point_reg=out['outputs_coord'], attns=out['attns'],
                              gt_points=gt_points, gt_points_labels=labels, return_mask=return_mask, pos_mask_thr=0.6,
                              neg_mask_thr=0.1, num_mask_point_gt=10, corr_size=21, obj_tau=0.85, pos_inds=pos_inds,
                              gt_index=gt_index)
    return res


def count_launches(fn):
    """Number of kernels launched by this repo's .so in one step (CUPTI via torch.profiler)."""
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        ours = ('linear_tcgen05', 'mhsa_fwd', 'attn_headmean', 'layernorm_f16', 'im2col16', 'assemble_tokens', 'rollout_', 'cam_',
                'ccl_', 'ms_', 'norm_', 'seed_proto', 'refine_', 'weighted_sum', 'fuse_', 'crop_', 'erode_down', 'filter_score',
                'merge_protos', 'part_', 'fill_u32', 'minmax_decode', 'ext_init')
        n_ours = n_all = 0
        for e in prof.events():
            if 'cuda' in str(e.device_type).lower() and e.name and not e.name.startswith('Memcpy') and not e.name.startswith('Memset'):
                n_all += 1
                if any(k in e.name for k in ours):
                    n_ours += 1
        return n_ours, n_all
    except Exception:
        return None, None


def run_ours(args):
    from attentionshift_b200 import parallel
    if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'INFO'):
        os.environ['NCCL_DEBUG'] = 'WARN'        # keep stdout to the single JSON line the driver parses
    rank, world, local = parallel.env_rank_world()
    # the host side of a rank is one launching thread: keep torch's CPU pool from oversubscribing the box when 8 ranks share it
    torch.set_num_threads(max(1, min(4, usable_cpus() // max(world, 1))))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    parallel.init('nccl', dev)
    cfg = dict(WORKLOAD)
    if args.small:
        cfg.update(batch=2, img=224, depth=2)
    from attentionshift_b200 import ops
    bb, head = build_models(cfg, dev)
    inputs = make_inputs(cfg, rank)
    img_host = inputs[0]
    img_dev = img_host.to(dev)

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize()

    # per-entry-point device timing: the library's timing slots are recorded on the launching stream around every C-ABI
    # call -- inside the backbone's CUDA graph too (captured during the first warm-up step), so the kernel times below
    # come from the timed region itself (its last step)
    if not os.environ.get('AS_BENCH_NO_TIMERS'):
        ops.TIMERS.enable()
    # the clock sampler starts BEFORE the warm-up: the first NVML queries of a process are slow and take a driver lock that
    # stalls kernel launches (seen as a 100+ ms hiccup in whichever loop ran first); its samples are reset when timing starts
    sampler = ClockSampler(local) if rank == 0 else None     # one per job: NVML calls serialise on a driver lock shared by all ranks
    if sampler is not None:
        sampler.start()
    for _ in range(max(args.warmup, 1)):
        ops.TIMERS.begin_step()
        one_step(bb, head, img_dev, inputs, False)
    torch.cuda.synchronize()

    # ---- device-resident timing (value)
    barrier()
    if sampler is not None:
        sampler.rows = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ops.TIMERS.begin_step()
        one_step(bb, head, img_dev, inputs, False)
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1) / args.steps
    fam = ops.TIMERS.summary() if ops.TIMERS.on else {}
    ops.TIMERS.disable()

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

    def e2e_loop(n):
        d2h = 0
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            res = one_step(bb, head, bufs[i % 2], inputs, True)
            freed[i % 2].record()
            d2h = sum(m.nbytes for m in res['pseudo_gt_masks'])
        return d2h

    for f in freed:
        f.record()
    e2e_loop(max(args.warmup, 1))          # untimed: first use of the pinned mask buffers / copy stream (cudaHostAlloc is slow)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    d2h = e2e_loop(args.steps)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    if sampler is not None:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    ms_dev, ms_e2e = parallel.max_over_ranks([ms_dev, ms_e2e], device=dev)     # the slowest rank defines the step
    n_ours, n_all = count_launches(lambda: one_step(bb, head, img_dev, inputs, False)) if rank == 0 else (None, None)

    if rank == 0:
        pk = peaks()
        B, T, C = cfg['batch'], 1 + (cfg['img'] // 16) ** 2 + cfg['n_point_tokens'], cfg['embed']
        N = (cfg['img'] // 16) ** 2
        att = fam.get('as_mhsa_fwd', {})
        flops_attn = 4.0 * T * T * C * B                      # SDPA part of F_attn, per launch (one layer, whole batch)
        roof = None
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'ncu_mhsa_traffic.json')      # dram bytes / launch from the committed ncu capture
        if os.path.exists(tpath) and not args.small:
            try:
                traffic = json.load(open(tpath)).get('dram_bytes_per_launch')
            except Exception:
                traffic = None
        if att.get('n'):
            ach = flops_attn / (att['ms'] / att['n'] * 1e-3) / 1e12
            roof = dict(kernel='mhsa_fwd2_kernel (tcgen05 flash attention, one launch = one layer x batch)', bound='tensor',
                        achieved=round(ach, 1), peak=pk['tf_sus'], unit='TFLOP/s', frac=round(ach / pk['tf_sus'], 4),
                        traffic=traffic, peak_source=pk['src'] + ' sustained bf16 GEMM', share_of_step=round(att['ms'] / ms_dev, 3))
        ms_name = next((k for k in ('as_mean_shift_fused', 'as_mean_shift_tc', 'as_mean_shift') if fam.get(k, {}).get('n')), 'as_mean_shift')
        msf = fam.get(ms_name, {})
        K = cfg['n_obj'] * cfg['seeds']
        b_alg = ((cfg['iters'] + 1) * N * C * 4 + K * N * 4 + 2 * K * C * 4) * B      # SURVEY 8d B_alg per image x images
        roof2 = None
        if msf.get('n'):
            ach2 = b_alg / (msf['ms'] / msf['n'] * 1e-3) / 1e9
            what = {'as_mean_shift_fused': 'one persistent cooperative kernel + the token split kernel',
                    'as_mean_shift_tc': '~50 launches', 'as_mean_shift': 'fp32 CUDA-core kernels'}[ms_name]
            traffic2 = None
            t2path = os.path.join(ROOT, 'profiles', 'ncu_msfused_traffic.json')      # dram bytes / launch from the committed ncu capture
            if ms_name == 'as_mean_shift_fused' and os.path.exists(t2path) and not args.small:
                try:
                    traffic2 = json.load(open(t2path)).get('dram_bytes_per_launch')
                except Exception:
                    traffic2 = None
            roof2 = dict(kernel='%s (whole on-device attention-shift loop, all images of the batch; %s)' % (ms_name, what), bound='hbm', achieved=round(ach2, 1),
                         peak=pk['hbm'], unit='GB/s', frac=round(ach2 / pk['hbm'], 4), traffic=traffic2, peak_source=pk['src'])
        line = dict(metric='images/sec at 1024^2 bs8 ViT-B attn-shift', value=round(world * B / (ms_dev * 1e-3), 2), unit='images/s',
                    n_gpus=world, steps=args.steps, warmup=max(args.warmup, 1), ms_per_step=round(ms_dev, 3), higher_is_better=True,
                    scaling='weak', vs_baseline=None, dtype='f16 operands / f32 accumulate (ViT GEMMs + attention), f32 (attention shift)',
                    data='synthetic (random-init ViT-B/16 weights, randn images, random GT points)',
                    config=dict(workload=cfg['name'], per_gpu_batch=B, fpn=False, l2='inputs larger than L2 (per-step working set >> 126 MB)',
                                backbone='one CUDA graph per forward (cuda_graph=True)', kernel_times='library timing slots, last step of the timed region',
                                attn_maps='last 7 layers (the ones seed_pseudo_gt reads)', small=bool(args.small)),
                    e2e=dict(value=round(world * B / (ms_e2e * 1e-3), 2), unit='images/s', ms_per_step=round(ms_e2e, 3),
                             h2d_bytes_per_step=int(img_host.nbytes), d2h_bytes_per_step=int(d2h)),
                    gpu_launches=n_ours, all_launches=n_all, clocks=sampler.summary(), roofline=roof, roofline_attnshift=roof2,
                    kernel_ms_per_step={k: round(v['ms'], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]['ms'])})
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline(cfg, steps=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------- CPU arm
def cpu_one_image(cfg, seed):
    """The reference algorithm (CPU oracle port, same torch ops as the reference) on ONE image of the workload."""
    import torch.nn.functional as F
    from attentionshift_b200.synthetic import structured_scene, vit_state_dict
    from oracle import attnshift as O
    from oracle import vit as V
    S = cfg['img']
    hp = S // 16
    g = torch.Generator().manual_seed(seed)
    sd = cpu_one_image.sd if hasattr(cpu_one_image, 'sd') else vit_state_dict(cfg['embed'], cfg['depth'], cfg['heads'], S,
                                                                            n_point_tokens=cfg['n_point_tokens'], seed=0)
    cpu_one_image.sd = sd
    img = torch.randn(1, 3, S, S, generator=g)
    t0 = time.time()
    with torch.no_grad():
        out = V.backbone_forward(img, sd, cfg['depth'], cfg['heads'], n_point_tokens=cfg['n_point_tokens'])
        t1 = time.time()
        rows = O.rollout_rows(out['attns'][-cfg['cam_layer']:], cfg['n_point_tokens'])
        # attention-shift stage on the structured scene: random-init features carry no instance structure
        sc = structured_scene(hp, hp, cfg['embed'], cfg['n_obj'], seed=seed, noise=0.5)
        up = F.interpolate(sc['cams_low'].reshape(-1, 1, hp, hp), (S, S), mode='bilinear').reshape(7, cfg['n_obj'], S, S)
        for l in range(7):
            for j in range(cfg['n_obj']):
                O.bbox_from_cam(up[l, j].clone(), sc['gt_points'][j], 0.2, 0.5, (S, S))
        O.attention_shift_image(up, sc['gt_index'], sc['rois'], sc['vit_feat'].clone(), sc['gt_points'], sc['gt_labels'],
                                mean_shift_times=cfg['iters'], n_points=20)
    t2 = time.time()
    return t2 - t0, t1 - t0, t2 - t1


def cpu_baseline(cfg, steps=1):
    torch.set_num_threads(usable_cpus())
    ts = [cpu_one_image(cfg, 100 + i) for i in range(steps)]
    tot = sum(t[0] for t in ts) / len(ts)
    return dict(value=round(1.0 / tot, 4), unit='images/s', cores=usable_cpus(), kind='port',
                sample=f'{steps} image(s) of the workload (1/{cfg["batch"]} batch): ViT forward {ts[0][1]:.1f}s + roll-out/attention-shift {ts[0][2]:.1f}s; '
                       'torch CPU fp32, all host threads; attention-shift stage on the structured synthetic scene (20 seeds as hard-coded in the reference)')


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cfg = dict(WORKLOAD)
    if args.small:
        cfg.update(batch=2, img=224, depth=2)
    torch.set_num_threads(usable_cpus())
    budget = 240.0
    t_all = time.time()
    done_w = 0
    for _ in range(min(args.warmup, 1)):
        cpu_one_image(cfg, 7)
        done_w += 1
    per = []
    for i in range(args.steps):
        if per and (time.time() - t_all) + per[-1] > budget:
            break
        per.append(cpu_one_image(cfg, 100 + i)[0])
    ms = 1e3 * sum(per) / len(per)
    v = round(1e3 / ms, 4)
    line = dict(impl='reference', metric='images/sec at 1024^2 bs8 ViT-B attn-shift', value=v, unit='images/s', n_gpus=args.gpus,
                steps=len(per), warmup=done_w, ms_per_step=round(ms, 1), higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic (same generator as the CUDA arm)',
                config=dict(workload=cfg['name'], per_gpu_batch=cfg['batch'], small=bool(args.small),
                            note='each step = ONE image of the batch (bounded sample); the reference algorithm is per-image, images/s is per host'),
                cpu_baseline=dict(value=v, unit='images/s', cores=usable_cpus(), kind='port',
                                  sample='one image per step; oracle port of the reference (the python reference cannot travel to the GPU box)'),
                e2e=dict(value=v, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
