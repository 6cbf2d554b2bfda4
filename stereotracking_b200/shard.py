"""Multi-GPU plumbing of the path: whole sequences per rank, no data-path collective.

The reference splits whole videos across ranks with ``np.array_split``
(mmtrack/datasets/samplers/video_sampler.py:50-69) and gathers per-sequence results
on the host after the last frame (mmtrack/evaluation/metrics/mot_drone_metrics.py
:336-345 ``all_gather_object(seq_info)``).  The same here: each rank owns a
contiguous block of sequences and their track state for their whole life; the only
communication is one ``gather_object`` of track rows at the end.  Host-side code,
backend-agnostic (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np
import torch.distributed as dist


def bind_to_gpu_numa(device_index):
    """Best effort: restrict this process to the CPUs of the NUMA node its GPU hangs off,
    before any pinned host buffer is allocated (first-touch places the pages there), so
    that with one process per GPU the host<->device copies of all ranks do not cross the
    socket interconnect.  Returns a small report dict; never raises."""
    import os
    try:
        bus = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get('CUDA_VISIBLE_DEVICES', '').split(',')[device_index])
                                                  if os.environ.get('CUDA_VISIBLE_DEVICES') else device_index)
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
        except Exception:
            import torch
            pr = torch.cuda.get_device_properties(device_index)
            bus = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        bus = bus.lower()
        if len(bus.split(':')[0]) == 8:                 # nvml prints an 8-digit domain
            bus = bus[4:]
        node = int(open(f'/sys/bus/pci/devices/{bus}/numa_node').read())
        if node < 0:
            return dict(bound=False, reason='numa_node=-1', bus=bus)
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return dict(bound=False, reason='no allowed cpu on the node', node=node, bus=bus)
        os.sched_setaffinity(0, allowed)
        return dict(bound=True, node=node, cpus=len(allowed), bus=bus)
    except Exception as e:                              # containers may hide sysfs / NVML
        return dict(bound=False, reason=f'{type(e).__name__}: {e}')


def shard_sequences(n_sequences, world_size, rank):
    """(first, count) of the contiguous block of sequences rank ``rank`` owns:
    ``np.array_split`` semantics (the first ``n % world`` ranks get one more),
    video_sampler.py:59-66.  Raises like the reference when there are fewer
    sequences than ranks (:55-57)."""
    if n_sequences < world_size:
        raise ValueError(f'only {n_sequences} videos loaded,but {world_size} gpus were given.')
    if not 0 <= rank < world_size:
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, extra = divmod(n_sequences, world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def tracks_to_rows(results, frame_id, first_sequence=0):
    """One lockstep frame of engine results (list of per-sequence dicts) -> float64
    rows ``[sequence, frame, id, x1, y1, x2, y2, score, depth, label]``."""
    rows = []
    for s, r in enumerate(results):
        n = len(r['instances_id'])
        if n == 0:
            continue
        block = np.empty((n, 10), np.float64)
        block[:, 0] = first_sequence + s
        block[:, 1] = frame_id
        block[:, 2] = r['instances_id']
        block[:, 3:7] = r['bboxes']
        block[:, 7] = r['scores']
        block[:, 8] = r['depth']
        block[:, 9] = r['labels']
        rows.append(block)
    return np.concatenate(rows) if rows else np.empty((0, 10), np.float64)


def gather_tracks(rows, dst=0, group=None):
    """Host-side gather of every rank's track rows on ``dst`` (None elsewhere), ordered
    by (sequence, frame, position); single-process runs return ``rows``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(rows, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    parts = [b for b in bucket if b is not None and len(b)]
    if not parts:
        return np.empty((0, 10), np.float64)
    out = np.concatenate(parts)
    return out[np.lexsort((out[:, 1], out[:, 0]))]          # stable: keeps the per-frame output order


def write_mot_txt(rows, sequence, path, depth_thr=None):
    """MOTChallenge txt of one sequence as MOTDroneMetrics writes it
    (mot_drone_metrics.py:195-207 depth-range filter ``depth_thr >= depth > 0`` and width / height
    as float32 differences of the float32 box, :231-236 line format; frame ids are 1-based).
    Pinned by tests/golden/ref_writers.npz."""
    sel = rows[rows[:, 0] == sequence]
    if depth_thr is not None:
        with np.errstate(invalid='ignore'):
            sel = sel[(sel[:, 8] > 0) & (sel[:, 8] <= depth_thr)]
    box = sel[:, 3:7].astype(np.float32)
    w, h = box[:, 2] - box[:, 0], box[:, 3] - box[:, 1]            # float32, as the reference's tensors
    with open(path, 'wt') as f:
        for k, r in enumerate(sel):
            f.write('%d,%d,%.3f,%.3f,%.3f,%.3f,%.3f,-1,-1,-1\n' %
                    (r[1] + 1, r[2], box[k, 0], box[k, 1], w[k], h[k], np.float32(r[7])))
    return len(sel)


def _csv_depth(v):
    """extract_depth's list entries (ocsort_disparity.py:148,:168): the int -1 for a box without a
    valid depth, else a numpy float32."""
    return -1 if v == -1 else np.float32(v)


def write_results_csv(rows, sequence, path, gt_depth=None):
    """``results.csv`` of one sequence as the reference's ``@save_prediction_results``
    decorator appends it per frame (mmtrack/utils/collect_results.py:25-37): header
    frame,id,label,tl_x,tl_y,br_x,br_y,depth,gt_depth,score; ``gt_depth`` falls back to
    the predicted depth as in ocsort_disparity.py:103-104.  Pinned byte for byte by
    tests/golden/ref_writers.npz (the reference's own decorator run on the same rows)."""
    import csv
    sel = rows[rows[:, 0] == sequence]
    with open(path, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['frame', 'id', 'label', 'tl_x', 'tl_y', 'br_x', 'br_y', 'depth', 'gt_depth', 'score'])
        for k, r in enumerate(sel):
            gd = r[8] if gt_depth is None else gt_depth[k]
            w.writerow([int(r[1]), int(r[2]), int(r[9]), np.float32(r[3]), np.float32(r[4]), np.float32(r[5]),
                        np.float32(r[6]), _csv_depth(r[8]), _csv_depth(gd), np.float32(r[7])])
    return len(sel)
