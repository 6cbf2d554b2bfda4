#!/usr/bin/env python
"""Benchmark of the depth-guided tracking path (BASELINE.json metric: tracked
frames/sec with sequences batched in lockstep; % of HBM roofline; host-CPU
reference beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one lockstep frame of every sequence of the workload (S frames).
Workload at every N: BASELINE.json configs[1] per GPU - 64 synthetic 1280x720
stereo sequences (left image is not an input of this path; raw u16 disparity +
YOLOX head maps are), up to 50 airborne objects per frame - so per-GPU work is
fixed as N grows (weak scaling; sequences are independent, no collective on the
data path).  One JSON line is printed by rank 0.
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

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOAD = 'C2'
SEQ_PER_GPU = 64
METRIC = 'tracked_frames_per_sec'
# dram__bytes_read.sum + dram__bytes_write.sum of one k1_disp2depth_flat launch at the default
# workload (64 x 1280x720), from the `ncu --set full` capture profiles/r01_k1_disp2depth_flat_full*
# (118.0 MB read + 188.0 MB written; the rest of the 241 MB written is still in L2 at kernel end)
K1_TRAFFIC_BYTES_C2_S64 = 117.98e6 + 188.03e6
UNIT = 'frames/s'
# engine capacities per workload (tracks / detections / post-threshold candidates per sequence)
CAPS = dict(C1=dict(max_dets=256, max_tracks=256, max_candidates=2048),
            C2=dict(max_dets=256, max_tracks=256, max_candidates=2048),
            C3=dict(max_dets=512, max_tracks=512, max_candidates=4096),
            C4=dict(max_dets=1024, max_tracks=512, max_candidates=4096))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--seq', type=int, default=SEQ_PER_GPU, help='sequences per GPU')
    ap.add_argument('--workload', default=WORKLOAD)
    ap.add_argument('--cpu-seqs', type=int, default=0, help='sequences of the CPU sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--groups', type=int, default=1, help='independent lockstep groups (handles) per GPU')
    ap.add_argument('--depth', type=int, default=5, help='frames in flight in the frame pipeline (1 = no overlap)')
    ap.add_argument('--depth-source', default='raw', choices=['raw', 'map'],
                    help="headline mode: 'raw' = K3 gathers from the raw u16 frame, no depth map is written; "
                         "'map' = K1 writes the float32 depth map and K3 gathers from it.  Both are measured.")
    ap.add_argument('--reps', type=int, default=5, help='repetitions of the K-step timed region (median reported)')
    ap.add_argument('--e2e-reps', type=int, default=3)
    ap.add_argument('--graph', type=int, default=1, help='replay each frame as a CUDA graph (0 = stream launches)')
    ap.add_argument('--parity-seqs', type=int, default=4, help='sequences replayed through the oracle after the timed regions')
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    return rank, world, local


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self, windows=None):
        """Summary of the samples received inside the (t0, t1) perf_counter windows (the timed
        regions; a sample is kept when it arrives up to 30 ms after a window closes)."""
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.1)
        self.proc.terminate()
        rows = [r for t, r in self.rows if windows is None or any(a <= t <= b + 0.03 for a, b in windows)]
        if not rows:                      # regions shorter than one sampling period: nearest samples
            rows = [r for _, r in self.rows][-3:]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), reasons=sorted(reasons))


# --------------------------------------------------------------------------- data
def generate(cfg, S, n_frames, seed, device, chunk=2):
    """Synthetic frames resident on `device`: raw [F,S,H,W] u16 and per-level head
    maps [F,S,C,h,w] f32 (stereotracking_b200/synthetic.py)."""
    from stereotracking_b200.synthetic import SyntheticStereo
    gen = SyntheticStereo(cfg, S, n_frames, seed=seed, device=device)
    parts = [gen.frames(f0, min(f0 + chunk, n_frames)) for f0 in range(0, n_frames, chunk)]
    raw = torch.cat([p['raw'] for p in parts])
    L = len(parts[0]['cls'])
    cls = [torch.cat([p['cls'][l] for p in parts]) for l in range(L)]
    box = [torch.cat([p['bbox'][l] for p in parts]) for l in range(L)]
    obj = [torch.cat([p['obj'][l] for p in parts]) for l in range(L)]
    return raw, cls, box, obj


# --------------------------------------------------------------------------- CPU arm
_CPU_DATA = {}


def _cpu_worker(s):
    """The oracle (CPU restatement of the reference tracker path) on one sequence."""
    from oracle.model import StereoTrackOracle
    torch.set_num_threads(1)
    d = _CPU_DATA
    o = StereoTrackOracle()
    F = d['raw'].shape[0]
    t0 = time.perf_counter()
    rows = 0
    for f in range(F):
        r = o.step(f, d['raw'][f, s], [c[f, s] for c in d['cls']], [c[f, s] for c in d['box']],
                   [c[f, s] for c in d['obj']])
        rows += len(r['instances_id'])
    return F, time.perf_counter() - t0, rows


def _cpu_warm(_):
    from oracle.model import StereoTrackOracle  # noqa: F401
    from oracle.lapjv import lapjv
    lapjv(np.zeros((1, 1), np.float32), 0.9)
    time.sleep(0.05)
    return 0


def cpu_reference(raw, cls, box, obj, n_seqs, cores):
    """frames/s of the oracle over `n_seqs` sequences, one process per core."""
    import multiprocessing as mp
    import warnings
    warnings.simplefilter('ignore')
    _CPU_DATA.update(raw=raw[:, :n_seqs].cpu().numpy(), cls=[c[:, :n_seqs].cpu().numpy() for c in cls],
                     box=[c[:, :n_seqs].cpu().numpy() for c in box], obj=[c[:, :n_seqs].cpu().numpy() for c in obj])
    cores = max(1, min(cores, n_seqs))
    if cores == 1:
        t0 = time.perf_counter()
        res = [_cpu_worker(s) for s in range(n_seqs)]
        wall = time.perf_counter() - t0
    else:
        with mp.get_context('fork').Pool(cores) as pool:
            pool.map(_cpu_warm, range(cores), chunksize=1)     # imports + library load, untimed
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, range(n_seqs), chunksize=1)
            wall = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    single = np.mean([r[0] / r[1] for r in res])
    return frames / wall, cores, float(single), frames, sum(r[2] for r in res)


# --------------------------------------------------------------------------- main
def main():
    args = parse()
    rank, world, local = dist_env()
    from stereotracking_b200.synthetic import CONFIGS
    cfg = CONFIGS[args.workload]
    S, K, W = args.seq, args.steps, max(args.warmup, 0)
    F = K + W
    config = dict(workload=f'{args.workload}: {S} synthetic {cfg.width}x{cfg.height} stereo sequences per GPU in '
                           f'lockstep, {cfg.k_min}-{cfg.k_max} objects/frame, {cfg.n_clutter} clutter priors, '
                           f'{int(cfg.invalid_prob * 100)}% invalid disparity',
                  sequences_per_gpu=S, frames_per_step=S * max(world, 1), priors_per_frame=cfg.num_priors,
                  tracker='OCSORTTracker_Disparity (configs/stereo_tracking/ocsort/yolox_s_mmyolo_mot_airdrone_disp.py)',
                  l2_policy='every step reads fresh frames (147 MB/step at C2) > 126 MB L2; no flush needed',
                  groups_per_gpu=args.groups, pipeline_depth=args.depth,
                  parallelism=f'{world} GPU(s) x {args.groups} lockstep group(s) x {S // max(args.groups, 1)} '
                              f'sequences, frame pipeline of depth {args.depth} (K1/K2/K3 of later frames overlap the '
                              f'tracker of earlier ones), no data-path collective')

    if args.impl == 'reference':
        if rank != 0:
            return
        # the reference's own CPU implementation of the path = the oracle port
        # (the reference is pure Python and cannot be installed here: mmengine/
        # mmcv/mmdet/mmyolo/lap are absent)
        cores = os.cpu_count() or 1
        n_seqs = args.cpu_seqs or max(2, min(S, cores))
        frames_per_seq = min(F, 24)
        dev = 'cuda' if torch.cuda.is_available() else 'cpu'
        raw, cls, box, obj = generate(cfg, n_seqs, frames_per_seq, 20240604 + 1000, dev)
        times = []
        for _ in range(max(1, min(K, 3))):
            fps, used, single, frames, rows = cpu_reference(raw, cls, box, obj, n_seqs, cores)
            times.append(fps)
        fps = float(np.median(times))
        sample = f'{n_seqs} sequences x {frames_per_seq} frames of the same workload, one process per core'
        line = dict(metric=METRIC, value=fps, unit=UNIT, n_gpus=args.gpus, steps=K, warmup=W,
                    ms_per_step=1e3 * S * world / fps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f32 (f64 Kalman/assignment)', data='synthetic', config=config, impl='reference',
                    cpu_baseline=dict(value=fps, unit=UNIT, cores=used, kind='port', sample=sample,
                                      single_core=single),
                    e2e=dict(value=fps, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), 'bench.py needs a GPU for --impl b200 (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    from stereotracking_b200.engine import EngineConfig, FrameBuffers, FrameInputs, LockstepGroups, StereoTrackEngine

    from stereotracking_b200 import shard
    # bind to the GPU's NUMA node BEFORE any pinned host allocation (first touch places the pages)
    numa = shard.bind_to_gpu_numa(local) if world > 1 else dict(bound=False, reason='single process')
    # whole sequences per rank (video_sampler.py:50-69): rank r owns sequences [first, first + S)
    first_seq, n_own = shard.shard_sequences(S * world, world, rank)
    assert n_own == S
    raw, cls, box, obj = generate(cfg, S, F, 20240604 + 1000 + first_seq, dev)
    level_hw = [c.shape[-2:] for c in cls]
    caps = CAPS.get(args.workload, CAPS['C2'])
    eng = StereoTrackEngine(EngineConfig(n_seq=S, **caps), device=dev)
    if os.environ.get('ST_K4_SKIP'):                       # diagnostics only (st_debug_set_skip bits)
        eng.lib.st_debug_set_skip(int(os.environ['ST_K4_SKIP']))
    eng.make_frame_io(raw.shape[1:], level_hw)      # allocates eng.depth_map
    stream = torch.cuda.current_stream(dev)
    modes = [args.depth_source, 'map' if args.depth_source == 'raw' else 'raw']
    groups_of = {}
    for mode in modes:
        g = LockstepGroups(S, args.groups, device=dev, depth=args.depth, depth_source=mode, **caps)
        g.bind(raw.shape[1:], level_hw)
        g.plan(raw, cls, box, obj)      # device addresses of every (frame, group) slice, computed once
        if args.graph:
            g.use_graphs(True)
        groups_of[mode] = g
    groups = groups_of[modes[0]]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run_stages(f0, f1, events, mode):
        """the same launches as the pipeline, one C call per stage on ONE stream so that each stage
        can be bracketed by events (launch gaps of an idle stream included)"""
        Hp, Wp = cfg.padded_hw
        for f in range(f0, f1):
            eng.frame_ids.fill_(f)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
            ev[0].record(stream)
            if mode == 'map':
                eng.disp2depth(raw[f], out=eng.depth_map)
            ev[1].record(stream)
            eng.decode_nms([c[f] for c in cls], [c[f] for c in box], [c[f] for c in obj],
                           (cfg.height, cfg.width)); ev[2].record(stream)
            if mode == 'map':
                eng.box_depth(eng.depth_map, eng.det_boxes, eng.det_counts, eng.det_depth, eng.det_scales,
                              eng.det_inflated)
            else:
                eng.box_depth_raw(raw[f], (Hp, Wp), eng.det_boxes, eng.det_counts, eng.det_depth, eng.det_scales,
                                  eng.det_inflated)
            ev[3].record(stream)
            eng.track_step(eng.frame_ids, eng.det_inflated, eng.det_scores, eng.det_labels, eng.det_scales,
                           eng.det_depth, eng.det_counts); ev[4].record(stream)
            eng.scale_boxes(eng.trk_boxes, eng.out_scales, eng.out_counts, True, eng.out_boxes)
            ev[5].record(stream)
            if mode == 'map':
                eng.box_depth(eng.depth_map, eng.out_boxes, eng.out_counts, eng.out_depth, eng.out_scale2, None)
            else:
                eng.box_depth_raw(raw[f], (Hp, Wp), eng.out_boxes, eng.out_counts, eng.out_depth, eng.out_scale2, None,
                                  prev_boxes=eng.det_boxes, prev_index=eng.out_src, prev_d=eng.det_depth)
            ev[6].record(stream)
            if events is not None:
                events.append(ev)

    # ---- resident-input timing (value): `reps` repetitions of [W warm-up steps, K timed steps] ----
    # every repetition replays frames 0 .. F-1 (frame 0 resets the trackers); the K timed steps are
    # bracketed by barrier + synchronize and CUDA events on the launching stream (fork / join)
    clock_windows = []
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    def timed_pass(g):
        for f in range(0, W):
            g.run_plan(f)
        g.join(stream)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        g.fork(stream)                 # the pipeline streams start after e0 ...
        for f in range(W, F):
            g.run_plan(f)              # one C call per group queues the frame
        issue = (time.perf_counter() - t0) * 1e3       # host time to issue the K steps (no sync inside)
        g.join(stream)                 # ... and e1 is recorded after all of them finished
        e1.record(stream)
        barrier()
        clock_windows.append((t0, time.perf_counter()))
        return e0.elapsed_time(e1), issue

    timing = {}
    for mode in modes:
        g = groups_of[mode]
        timed_pass(g)                                   # one untimed repetition (graph instantiation, caches)
        reps = [timed_pass(g) for _ in range(max(1, args.reps))]
        ms = torch.tensor([r[0] for r in reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)   # per repetition: the slowest rank
        timing[mode] = dict(ms=ms.cpu().numpy(), issue=float(np.median([r[1] for r in reps])))
    status = groups.status()
    dbg = np.concatenate([e.debug_words() for e in groups.engines])
    last = {mode: groups_of[mode].results() for mode in modes}
    tracked_rows = int(sum(len(r['instances_id']) for r in last[modes[0]]))
    # both depth sources must give the same tracks (bit for bit) - checked on the last frame of all sequences
    modes_identical = all(np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == 'f')
                          for a, b in zip(last[modes[0]], last[modes[1]]) for k in a)

    # ---- per-kernel timing (serial single-stream passes over the same frames) -------------------
    stage_avg = {}
    for mode in modes:
        evs = []
        run_stages(0, W, None, mode)
        barrier()
        run_stages(W, F, evs, mode)
        barrier()
        stage_avg[mode] = np.array([[ev[i].elapsed_time(ev[i + 1]) for i in range(6)] for ev in evs]).mean(0)
    # K1 alone, back to back over the timed frames (fresh 118 MB read per launch, 241 MB written to
    # alternating maps): the launch queue stays full, so the average is the kernel's own duration;
    # the per-stage events above also contain the launch gap of an idle stream (~8 us per stage)
    k1_out = [eng.depth_map, torch.empty_like(eng.depth_map)]
    for f in range(min(W, 3)):
        eng.disp2depth(raw[f], out=k1_out[f % 2])
    barrier()
    k1_reps = []
    for _ in range(3):
        k0, k1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for f in range(W, F):
            eng.disp2depth(raw[f], out=k1_out[f % 2])
        k1e.record(stream)
        barrier()
        k1_reps.append(k0.elapsed_time(k1e) / K)
    k1_ms = float(np.median(k1_reps))
    del k1_out

    # ---- end to end through the public call with HOST buffers ---------------------
    # Per frame: ONE pinned host -> device copy of the frame's inputs (frame ids, raw disparity and
    # all head maps are one staging record, engine.FrameInputs) on the copy stream, the frame on the
    # frame pipeline, ONE device -> pinned host copy of its tracks (FrameBuffers.out_block) on the
    # read-back stream.  `depth` frames are in flight; the host consumes the tracks of frame f
    # while frame f + depth - 1 is being copied in.
    N = args.depth
    H_, W_ = cfg.height, cfg.width
    host = [FrameInputs(S, H_, W_, level_hw, device='cpu', pin=True).fill(
        f, raw[f].cpu(), [c[f].cpu() for c in cls], [c[f].cpu() for c in box], [c[f].cpu() for c in obj])
        for f in range(F)]
    d_in = [FrameInputs(S, H_, W_, level_hw, device=dev) for _ in range(N)]
    pipe = groups.pipes[0]
    assert len(groups.pipes) == 1 or not args.graph, 'one lockstep group per GPU in the end-to-end region'
    h_out = [[torch.empty_like(b.out_block, device='cpu').pin_memory() for b in p.slots] for p in groups.pipes]
    h2d = host[0].nbytes
    d2h = sum(p.slots[0].out_block.numel() for p in groups.pipes)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    G = len(groups.pipes)
    mk = lambda: torch.cuda.Event()
    ev_in = [mk() for _ in range(N)]
    ev_done = [[mk() for _ in range(G)] for _ in range(N)]
    ev_out = [mk() for _ in range(N)]
    for e in ev_in + ev_out + [x for row in ev_done for x in row]:
        e.record(stream)                   # materialise the CUDA events
    torch.cuda.synchronize(dev)

    def e2e_steps(f0, f1, collect=None):
        """frames f0 .. f1-1 from host staging records to host track rows; ``collect``: list that
        receives (frame, results of the sequences in parity_seqs) for every frame"""
        consumed = 0

        def consume(slot, f):
            nonlocal consumed
            ev_out[slot].synchronize()          # the tracks of frame f are on the host
            for g_, p in enumerate(groups.pipes):
                v = FrameBuffers.unpack(h_out[g_][slot], p.slots[slot].out_layout)
                consumed += int(v['out_counts'].sum())
            if collect is not None:
                res = []
                for g_, p in enumerate(groups.pipes):
                    res.extend(p.slots[slot].results(h_out[g_][slot].clone()))
                collect.append((f, [res[s] for s in parity_seqs]))

        for k, f in enumerate(range(f0, f1)):
            slot = groups.pipes[0].submitted % N
            if k >= N:
                consume(slot, f - N)            # the caller has read the tracks of frame f - N
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_out[slot])   # this slot's previous frame left the device
                d = d_in[slot]
                d.block.copy_(host[f].block, non_blocking=True)       # one cudaMemcpyAsync per frame
                ev_in[slot].record(s_in)
            for g_, (p, a, sz) in enumerate(zip(groups.pipes, groups.offsets, groups.sizes)):
                p.submit(d.frame_ids[a:a + sz], d.raw[a:a + sz], [c[a:a + sz] for c in d.cls],
                         [c[a:a + sz] for c in d.bbox], [c[a:a + sz] for c in d.obj],
                         ready_event=ev_in[slot], done_event=ev_done[slot][g_])
            with torch.cuda.stream(s_out):
                for g_, p in enumerate(groups.pipes):
                    s_out.wait_event(ev_done[slot][g_])
                    h_out[g_][slot].copy_(p.slots[slot].out_block, non_blocking=True)   # one copy per frame
                ev_out[slot].record(s_out)
        n_frames = f1 - f0
        for k in range(max(n_frames - N, 0), n_frames):
            consume((groups.pipes[0].submitted - n_frames + k) % N, f0 + k)
        return consumed

    parity_seqs = sorted({int(round(i * (S - 1) / max(args.parity_seqs - 1, 1))) for i in range(args.parity_seqs)}) \
        if args.parity_seqs > 0 else []
    e2e_steps(0, F)                             # untimed: graph instantiation for the staging addresses, caches
    e2e_ms_reps = []
    for _ in range(max(1, args.e2e_reps)):
        e2e_steps(0, W)
        barrier()
        t0 = time.perf_counter()
        e2e_steps(W, F)
        torch.cuda.synchronize(dev)
        mine = (time.perf_counter() - t0) * 1e3
        barrier()
        clock_windows.append((t0, time.perf_counter()))
        e2e_ms_reps.append(mine)
    # the host -> device copies alone, all ranks at once (what PCIe / host memory give this many ranks):
    # the ceiling the end-to-end number is measured against
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_in):
        c0.record(s_in)
        for f in range(W, F):
            d_in[f % N].block.copy_(host[f].block, non_blocking=True)
        c1.record(s_in)
    t0 = time.perf_counter()
    barrier()
    clock_windows.append((t0, time.perf_counter()))
    copy_ms = c0.elapsed_time(c1) / K
    clocks = sampler.stop(clock_windows) if rank == 0 else None      # samples of all timed regions

    # ---- parity of the timed workload: the oracle replays some of this rank's sequences ----------
    parity = None
    if parity_seqs:
        got = []
        e2e_steps(0, F, collect=got)
        got.sort(key=lambda t: t[0])
        from oracle.model import StereoTrackOracle
        mism = rows = 0
        first_bad = None
        for k, s_ in enumerate(parity_seqs):
            o = StereoTrackOracle()
            for f, res in got:
                ref = o.step(f, host[f].raw[s_].numpy(), [c[s_].numpy() for c in host[f].cls],
                             [c[s_].numpy() for c in host[f].bbox], [c[s_].numpy() for c in host[f].obj])
                r = res[k]
                ok = (np.array_equal(r['instances_id'], ref['instances_id']) and
                      np.array_equal(r['labels'], ref['labels']) and
                      np.allclose(r['bboxes'], ref['bboxes'], rtol=1e-5, atol=1e-4) and
                      np.allclose(r['scores'], ref['scores'], rtol=1e-5) and
                      np.allclose(r['depth'], ref['depth'], rtol=1e-5, equal_nan=True))
                rows += len(ref['instances_id'])
                if not ok:
                    mism += 1
                    first_bad = first_bad or dict(sequence=first_seq + s_, frame=f)
        parity = dict(sequences=[first_seq + s_ for s_ in parity_seqs], frames=F, rows=rows, mismatches=mism,
                      first_mismatch=first_bad, depth_sources_identical=bool(modes_identical),
                      checked='ids / labels bit-exact, boxes / scores / depths 1e-5 relative, every frame, against '
                              'the oracle (CPU restatement of the reference) on the same inputs')

    # ---- host-side gather of the last frame's tracks (outside every timed region) -------
    rows = shard.tracks_to_rows(groups.results(), F - 1, first_sequence=first_seq)
    gathered = shard.gather_tracks(rows, dst=0)

    # ---- reduce over ranks ------------------------------------------------------------
    e2e_t = torch.tensor(e2e_ms_reps + [copy_ms * K], dtype=torch.float64, device=dev)   # last entry: copies alone
    per_rank_e2e = [e2e_t.cpu().numpy()]
    per_rank_numa = [numa]
    bad = torch.tensor([0 if parity is None else parity['mismatches'] + (0 if modes_identical else 1)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        all_e2e = [torch.zeros_like(e2e_t) for _ in range(world)]
        dist.all_gather(all_e2e, e2e_t)
        per_rank_e2e = [t.cpu().numpy() for t in all_e2e]
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)          # per repetition: the slowest rank
        dist.all_reduce(bad, op=dist.ReduceOp.SUM)
        per_rank_numa = [None] * world
        dist.all_gather_object(per_rank_numa, numa)
    e2e_all = e2e_t.cpu().numpy()
    e2e_t, copy_only_ms = e2e_all[:-1], float(e2e_all[-1]) / K            # max over ranks of the copy-only pass
    e2e_ms = float(np.median(e2e_t))
    frames = S * world * K
    ms_all = timing[modes[0]]['ms']
    ms_total = float(np.median(ms_all))
    value = frames / (ms_total * 1e-3)
    e2e_value = frames / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        Hp, Wp = cfg.padded_hw
        k1_bytes = S * (2 * cfg.height * cfg.width + 4 * Hp * Wp)          # u16 read + f32 depth write
        k1_gbs = k1_bytes / (k1_ms * 1e-3) / 1e9
        P = cfg.num_priors
        k2_bytes = S * P * 4                                               # objectness logit of every prior (class logits lazily)
        names = ['k1_disp2depth', 'k2_decode_nms(2 kernels)', 'k3_box_depth(dets)', 'k4_track_step',
                 'scale_boxes', 'k3_box_depth(tracks)']
        kernels = {}
        for mode in modes:
            sa = stage_avg[mode]
            kernels[mode] = {n: dict(ms=float(sa[i]), share=float(sa[i] / sa.sum())) for i, n in enumerate(names)
                             if not (mode == 'raw' and i == 0)}
        kernels['map']['k1_disp2depth'].update(algorithmic_bytes=k1_bytes, ms_back_to_back=k1_ms, achieved_gbs=k1_gbs,
                                               frac=k1_gbs / peak)
        for mode in modes:
            kernels[mode]['k2_decode_nms(2 kernels)'].update(
                algorithmic_bytes=k2_bytes, achieved_gbs=k2_bytes / (stage_avg[mode][1] * 1e-3) / 1e9)
        # bytes one step of the pipeline moves through HBM, by depth source (algorithmic: the raw
        # frame or the map, the objectness plane of the head maps; box / track rows are < 1 MB)
        step_bytes = dict(map=k1_bytes + k2_bytes, raw=k2_bytes,
                          note='raw: K3 gathers ~2 bytes per box pixel from the u16 frame (about 4 MB per step at '
                               'this workload) instead of K1 streaming the frame into a float32 map')
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n_seqs = args.cpu_seqs or max(2, min(S, cores))
            fsub = min(F, 24)
            fps, used, single, nfr, rows_ = cpu_reference(raw[:fsub], [c[:fsub] for c in cls], [c[:fsub] for c in box],
                                                          [c[:fsub] for c in obj], n_seqs, cores)
            cpu = dict(value=fps, unit=UNIT, cores=used, kind='port', single_core=single,
                       sample=f'first {n_seqs} sequences x {fsub} frames of this workload, oracle (NumPy/C port of '
                              f'the reference tracker path), one process per core')
        trk_sms, other_sms = groups.pipes[0].partition()
        config.update(depth_source=modes[0], cuda_graph=bool(args.graph), value_repetitions=int(len(ms_all)),
                      e2e_repetitions=int(len(e2e_ms_reps)),
                      sm_partition=dict(tracker_sms=trk_sms, other_sms=other_sms) if trk_sms else None)
        other = modes[1]
        oms = float(np.median(timing[other]['ms']))
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total / K,
            higher_is_better=True, scaling='weak', vs_baseline=None,
            dtype='f32 (u16 disparity in, f64 Kalman state and assignment)', data='synthetic', config=config,
            value_stats=dict(median_ms_per_step=ms_total / K, min_ms_per_step=float(ms_all.min()) / K,
                             max_ms_per_step=float(ms_all.max()) / K, repetitions=int(len(ms_all)),
                             note='median over repetitions of the K-step region, each the max over ranks'),
            modes={modes[0]: dict(value=value, ms_per_step=ms_total / K, host_issue_ms_per_step=timing[modes[0]]['issue'] / K,
                                  hbm_bytes_per_step=step_bytes[modes[0]]),
                   other: dict(value=frames / (oms * 1e-3), ms_per_step=oms / K,
                               host_issue_ms_per_step=timing[other]['issue'] / K,
                               hbm_bytes_per_step=step_bytes[other],
                               min_ms_per_step=float(timing[other]['ms'].min()) / K,
                               max_ms_per_step=float(timing[other]['ms'].max()) / K),
                   'note': step_bytes['note']},
            roofline=dict(kernel='k1_disp2depth_flat', bound='hbm', achieved=k1_gbs, peak=peak, unit='GB/s',
                          frac=k1_gbs / peak,
                          traffic=K1_TRAFFIC_BYTES_C2_S64 if (args.workload == 'C2' and S == 64) else None,
                          traffic_unit='bytes per launch (ncu --set full, profiles/r01_k1_disp2depth_flat_full_metrics.csv)',
                          algorithmic_bytes_per_launch=k1_bytes, peak_source=peak_src,
                          whole_step_map_mode=dict(bytes=step_bytes['map'], ms=(ms_total if modes[0] == 'map' else oms) / K,
                                                   frac=step_bytes['map'] / ((ms_total if modes[0] == 'map' else oms) / K * 1e-3) / 1e9 / peak),
                          note='HBM-bound kernel of the path (north_star kernel 1): average duration of K launches '
                               'issued back to back on one stream over the timed frames (CUDA events, median of 3 '
                               'passes).  In depth_source=map it is a stage of every step; in depth_source=raw the '
                               'step no longer writes the map (K3 converts the box pixels on the fly) and the step '
                               'is latency-bound: "modes" gives both, "kernels" every stage of a serial pass'),
            kernels=kernels,
            cpu_baseline=cpu,
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                     ms_per_step=e2e_ms / K, h2d_copies_per_step=1, d2h_copies_per_step=G,
                     repetitions_ms_per_step=[float(v) / K for v in e2e_t],
                     h2d_copy_only_ms_per_step=copy_only_ms,
                     frac_of_copy_ceiling=copy_only_ms / (e2e_ms / K),
                     note='h2d_copy_only: the same pinned staging records copied back to back by all ranks at once, '
                          'nothing else running (slowest rank) - the host side ceiling of this box for this many ranks',
                     per_rank=[dict(rank=r, ms_per_step=float(np.median(v[:-1])) / K,
                                    h2d_gbs=h2d / (float(np.median(v[:-1])) / K * 1e-3) / 1e9,
                                    h2d_copy_only_gbs=h2d / (float(v[-1]) / K * 1e-3) / 1e9, numa=per_rank_numa[r])
                               for r, v in enumerate(per_rank_e2e)]),
            parity=parity,
            gpu_launches=int(K * groups.kernels_per_frame), clocks=clocks,
            host_issue_ms_per_step=timing[modes[0]]['issue'] / K,
            k4_phase_cycles_median=dict(zip(['filter_predict', 'stage1', 'stage2', 'stage3', 'smooth_order', 'update'],
                                            [int(x) for x in np.median(dbg[:, 8:14], axis=0)])),
            k4_phase_cycles_slowest_sequence=dict(zip(
                ['filter_predict', 'stage1', 'stage2', 'stage3', 'smooth_order', 'update'],
                [int(x) for x in dbg[int(np.argmax(dbg[:, 8:14].sum(1))), 8:14]])),
            assignments=dict(solved=int(dbg[:, 5].sum()), exact_fallback=int(dbg[:, 6].sum())),
            status_bits=int(status.max()), tracked_rows_last_step=tracked_rows,
            gathered_rows_all_ranks=int(len(gathered)))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if float(bad[0]) != 0:
        sys.stderr.write('PARITY MISMATCH: the timed workload differs from the oracle (see "parity")\n')
        sys.exit(3)


if __name__ == '__main__':
    main()
