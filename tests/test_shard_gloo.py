"""N > 1 host path on the CPU: whole sequences per rank, one gather at the end
(world_size 2, gloo).  The per-rank tracker here is the oracle - the point of the
test is the sharding / gather / writer logic around it, which is what runs under
torchrun on the GPU box (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from stereotracking_b200 import shard  # noqa: E402
from stereotracking_b200.synthetic import SceneConfig, SyntheticStereo  # noqa: E402

N_SEQ, N_FRAMES = 5, 6
CFG = SceneConfig(96, 128, 2, 5, n_clutter=5)


def _rows_for_block(first, count):
    """Track rows of sequences [first, first+count) from the CPU oracle."""
    from oracle.model import StereoTrackOracle
    fr = SyntheticStereo(CFG, N_SEQ, N_FRAMES, seed=41).frames(0, N_FRAMES)
    oracles = [StereoTrackOracle() for _ in range(count)]
    rows = []
    for f in range(N_FRAMES):
        res = []
        for k, o in enumerate(oracles):
            s = first + k
            res.append(o.step(f, fr['raw'][f, s].numpy(), [c[f, s].numpy() for c in fr['cls']],
                              [c[f, s].numpy() for c in fr['bbox']], [c[f, s].numpy() for c in fr['obj']]))
        rows.append(shard.tracks_to_rows(res, f, first_sequence=first))
    return np.concatenate(rows)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        first, count = shard.shard_sequences(N_SEQ, world, rank)
        rows = _rows_for_block(first, count)
        gathered = shard.gather_tracks(rows, dst=0)
        if rank == 0:
            np.save(os.path.join(out_dir, 'gathered.npy'), gathered)
        else:
            assert gathered is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_sequences_matches_array_split():
    for n in (2, 5, 64, 512, 513):
        for world in (1, 2, 4, 8):
            if n < world:
                continue
            chunks = np.array_split(np.arange(n), world)
            for r in range(world):
                first, count = shard.shard_sequences(n, world, r)
                assert (first, count) == (int(chunks[r][0]), len(chunks[r]))
    with pytest.raises(ValueError):
        shard.shard_sequences(1, 2, 0)          # fewer videos than ranks (video_sampler.py:55-57)


@pytest.mark.timeout(300)
def test_two_ranks_equal_one_rank(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    gathered = np.load(tmp_path / 'gathered.npy')
    single = _rows_for_block(0, N_SEQ)
    single = single[np.lexsort((single[:, 1], single[:, 0]))]
    assert len(gathered) > 20
    np.testing.assert_array_equal(gathered, single)
    # MOTChallenge writer: 1-based frames, xywh, depth-range filter
    n = shard.write_mot_txt(gathered, 1, tmp_path / 'seq1.txt', depth_thr=80)
    lines = open(tmp_path / 'seq1.txt').read().strip().splitlines()
    sel = gathered[(gathered[:, 0] == 1) & (gathered[:, 8] > 0) & (gathered[:, 8] <= 80)]
    assert n == len(lines) == len(sel)
    if n:
        f0 = lines[0].split(',')
        assert int(f0[0]) == int(sel[0, 1]) + 1 and int(f0[1]) == int(sel[0, 2]) and f0[-3:] == ['-1', '-1', '-1']
        assert abs(float(f0[4]) - (sel[0, 5] - sel[0, 3])) < 1e-3


def test_gather_is_identity_without_process_group():
    rows = np.arange(20, dtype=np.float64).reshape(2, 10)
    assert shard.gather_tracks(rows) is rows


def test_results_csv_has_the_reference_columns(tmp_path):
    rows = np.array([[3, 0, 7, 10.5, 20.25, 30, 40, 0.9, 12.5, 0],
                     [3, 1, 7, 11.5, 21.25, 31, 41, 0.8, np.nan, 0],
                     [4, 0, 1, 0, 0, 5, 5, 0.5, 3.0, 0]], np.float64)
    assert shard.write_results_csv(rows, 3, tmp_path / 'results.csv') == 2
    lines = open(tmp_path / 'results.csv').read().strip().splitlines()
    assert lines[0] == 'frame,id,label,tl_x,tl_y,br_x,br_y,depth,gt_depth,score'   # collect_results.py:36
    assert lines[1] == '0,7,0,10.5,20.25,30.0,40.0,12.5,12.5,0.9'
    assert lines[2].split(',')[7] == 'nan' and len(lines) == 3
