"""Parity at the BASELINE.json configurations beyond C1/C2-small: C3 (1920x1080, up to 200
tracks per sequence) and C4 (30 % invalid disparity, heavy ego-motion, ~1000 candidates per
frame before NMS) against the oracle on sizes it finishes in seconds, and size-independent
properties at the full C2 bench size (64 sequences x 1280x720)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mods():
    import torch
    from types import SimpleNamespace
    from oracle.model import StereoTrackOracle
    from stereotracking_b200.engine import EngineConfig, FramePipeline, StereoTrackEngine
    from stereotracking_b200.synthetic import CONFIGS, SceneConfig, SyntheticStereo
    return SimpleNamespace(torch=torch, StereoTrackOracle=StereoTrackOracle, EngineConfig=EngineConfig,
                           FramePipeline=FramePipeline, StereoTrackEngine=StereoTrackEngine, CONFIGS=CONFIGS,
                           SceneConfig=SceneConfig, SyntheticStereo=SyntheticStereo)


def _pipeline_vs_oracle(m, cfg, S, F, seed, **eng_kw):
    """Frames go through the frame pipeline (st_pipeline_submit, 5 streams, K3 reuse); the
    oracle replays the same inputs sequence by sequence."""
    torch = m.torch
    fr = m.SyntheticStereo(cfg, S, F, seed=seed).frames(0, F)
    raw = fr['raw'].cuda().contiguous()
    heads = [[c.cuda().contiguous() for c in fr[k]] for k in ('cls', 'bbox', 'obj')]
    fid = torch.arange(F, dtype=torch.int32, device='cuda')[:, None].repeat(1, S).contiguous()
    eng = m.StereoTrackEngine(m.EngineConfig(n_seq=S, **eng_kw))
    pipe = m.FramePipeline(eng, (S, cfg.height, cfg.width), [c.shape[-2:] for c in heads[0]], depth=F)
    for f in range(F):
        pipe.submit(fid[f], raw[f], [c[f] for c in heads[0]], [c[f] for c in heads[1]], [c[f] for c in heads[2]])
    pipe.join()
    torch.cuda.synchronize()
    assert eng.status().max() == 0
    oracles = [m.StereoTrackOracle() for _ in range(S)]
    rows = max_tracks = max_cand = 0
    for f in range(F):
        got = pipe.slots[f].results()
        for s in range(S):
            ref = oracles[s].step(f, fr['raw'][f, s].numpy(), [c[f, s].numpy() for c in fr['cls']],
                                  [c[f, s].numpy() for c in fr['bbox']], [c[f, s].numpy() for c in fr['obj']])
            np.testing.assert_array_equal(got[s]['instances_id'], ref['instances_id'], err_msg=f'frame {f} seq {s}')
            np.testing.assert_array_equal(got[s]['labels'], ref['labels'])
            np.testing.assert_allclose(got[s]['bboxes'], ref['bboxes'], rtol=1e-5, atol=1e-4)
            np.testing.assert_allclose(got[s]['scores'], ref['scores'], rtol=2e-7)
            np.testing.assert_allclose(got[s]['depth'], ref['depth'], rtol=1e-5, equal_nan=True)
            rows += len(ref['instances_id'])
            max_tracks = max(max_tracks, len(oracles[s].tracker.export_state()['ids']))
    return rows, max_tracks


def test_c3_1080p_many_tracks_matches_oracle(mods):
    rows, max_tracks = _pipeline_vs_oracle(mods, mods.CONFIGS['C3'], S=2, F=10, seed=51, max_dets=512,
                                           max_tracks=512, max_candidates=4096)
    assert rows > 1500 and max_tracks > 100, (rows, max_tracks)


def test_c4_stress_thousand_candidates_matches_oracle(mods):
    rows, _ = _pipeline_vs_oracle(mods, mods.CONFIGS['C4'], S=2, F=12, seed=52, max_dets=1024, max_tracks=512,
                                  max_candidates=4096)
    assert rows > 300, rows


def test_full_c2_size_properties(mods):
    """64 x 1280x720 in lockstep (the bench workload): K1 bit-exact against the float32
    expression on every pixel; the frame pipeline equals the single-stream path on all 64
    sequences; ids are unique per frame and never reused; a batch equals its sequences."""
    m, torch = mods, mods.torch
    cfg, S, F = m.CONFIGS['C2'], 64, 6
    gen = m.SyntheticStereo(cfg, S, F, seed=20240604 + 1000, device='cuda')
    fr = gen.frames(0, F)
    raw, heads = fr['raw'], [fr['cls'], fr['bbox'], fr['obj']]
    eng = m.StereoTrackEngine(m.EngineConfig(n_seq=S))
    # K1 at full size
    depth, _, _ = eng.disp2depth(raw[0])
    r = raw[0].cpu().numpy().astype(np.float32)
    disp = np.where(r == 65535, np.float32(0), r * np.float32(0.0625)).astype(np.float32)
    ref = np.full((S,) + cfg.padded_hw, np.float32(1.0) / np.float32(1e-6) * np.float32(160.0), np.float32)
    ref[:, :cfg.height, :cfg.width] = (np.float32(1.0) / (disp + np.float32(1e-6))) * np.float32(160.0)
    assert np.array_equal(depth.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    # sequential path
    want = []
    for f in range(F):
        eng.frame_ids.fill_(f)
        eng.step(raw[f], [c[f] for c in heads[0]], [c[f] for c in heads[1]], [c[f] for c in heads[2]])
        want.append(eng.results())
    # pipelined path
    eng2 = m.StereoTrackEngine(m.EngineConfig(n_seq=S))
    pipe = m.FramePipeline(eng2, (S, cfg.height, cfg.width), [c.shape[-2:] for c in heads[0]], depth=F)
    fid = torch.arange(F, dtype=torch.int32, device='cuda')[:, None].repeat(1, S).contiguous()
    for f in range(F):
        pipe.submit(fid[f], raw[f], [c[f] for c in heads[0]], [c[f] for c in heads[1]], [c[f] for c in heads[2]])
    pipe.join()
    torch.cuda.synchronize()
    seen = [set() for _ in range(S)]
    high = [-1] * S
    total = 0
    for f in range(F):
        got = pipe.slots[f].results()
        for s in range(S):
            for k in want[f][s]:
                np.testing.assert_array_equal(got[s][k], want[f][s][k], err_msg=f'frame {f} seq {s} {k}')
            ids = got[s]['instances_id']
            assert len(set(ids.tolist())) == len(ids)                      # unique within a frame
            new = [i for i in ids.tolist() if i not in seen[s]]
            assert all(i > high[s] for i in new)                            # fresh ids are never reused
            seen[s].update(new)
            high[s] = max([high[s]] + new)
            total += len(ids)
    assert total > 5000 and eng.status().max() == 0 and eng2.status().max() == 0
    # a batch equals its sequences: replay sequence 17 alone
    one = m.StereoTrackEngine(m.EngineConfig(n_seq=1))
    for f in range(F):
        one.frame_ids.fill_(f)
        one.step(raw[f, 17:18].contiguous(), [c[f, 17:18].contiguous() for c in heads[0]],
                 [c[f, 17:18].contiguous() for c in heads[1]], [c[f, 17:18].contiguous() for c in heads[2]])
        r1 = one.results()[0]
        for k in r1:
            np.testing.assert_array_equal(r1[k], want[f][17][k])


def _state_equal(st, rs):
    """Tracker state of one sequence (st_export_state) against the oracle's: ids, flags and
    counters bit-exact, Kalman state and velocities within 1e-5 relative (north_star)."""
    np.testing.assert_array_equal(st['ids'], rs['ids'])
    np.testing.assert_array_equal(st['tentative'], rs['tentative'])
    np.testing.assert_array_equal(st['tracked'], rs['tracked'])
    np.testing.assert_array_equal(st['hits'], rs['hits'])
    np.testing.assert_array_equal(st['last_frame'], rs['last_frame'])
    assert st['num_tracks'] == rs['num_tracks']
    np.testing.assert_allclose(st['mean'], rs['mean'], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(st['cov'], rs['cov'], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(st['velocity'], rs['velocity'], rtol=1e-5, atol=1e-6)


def _rows_equal(got, ref, tag):
    np.testing.assert_array_equal(got['instances_id'], ref['instances_id'], err_msg=tag)
    np.testing.assert_array_equal(got['labels'], ref['labels'], err_msg=tag)
    np.testing.assert_allclose(got['bboxes'], ref['bboxes'], rtol=1e-5, atol=1e-4, err_msg=tag)
    np.testing.assert_allclose(got['scores'], ref['scores'], rtol=2e-7, err_msg=tag)
    np.testing.assert_allclose(got['scales'], ref['scales'], rtol=1e-5, equal_nan=True, err_msg=tag)
    np.testing.assert_allclose(got['depth'], ref['depth'], rtol=1e-5, equal_nan=True, err_msg=tag)


@pytest.mark.parametrize('depth_source', ['map', 'raw'])
def test_bench_workload_c2_s64_matches_oracle(mods, depth_source):
    """BASELINE.json configs[1] exactly as bench.py runs it - 64 x 1280x720 sequences in lockstep,
    bench.py's seed and generator, the frame pipeline with a ring of 5 buffer sets - against the
    oracle: 8 of the 64 sequences x 60 frames (age-based eviction at 30 frames, re-found tracks
    and a few hundred ids per sequence fall inside), every output row; tracker state at the end."""
    m, torch = mods, mods.torch
    cfg, S, F, depth = m.CONFIGS['C2'], 64, 60, 5
    check = [0, 9, 17, 26, 35, 44, 53, 63]
    gen = m.SyntheticStereo(cfg, S, F, seed=20240604 + 1000, device='cuda')
    parts = [gen.frames(f0, f0 + 2) for f0 in range(0, F, 2)]
    raw = torch.cat([p['raw'] for p in parts])
    heads = [[torch.cat([p[k][l] for p in parts]) for l in range(3)] for k in ('cls', 'bbox', 'obj')]
    del parts
    host_raw = raw.view(torch.int16)[:, check].cpu().numpy().view(np.uint16)      # no uint16 indexing kernel in torch
    host_heads = [[c[:, check].cpu().numpy() for c in hk] for hk in heads]
    fid = torch.arange(F, dtype=torch.int32, device='cuda')[:, None].repeat(1, S).contiguous()
    eng = m.StereoTrackEngine(m.EngineConfig(n_seq=S, max_dets=256, max_tracks=256, max_candidates=2048))
    pipe = m.FramePipeline(eng, (S, cfg.height, cfg.width), [c.shape[-2:] for c in heads[0]], depth=depth,
                           depth_source=depth_source)
    done = [torch.cuda.Event() for _ in range(depth)]
    for e in done:
        e.record()
    got = []

    def collect(f):
        done[f % depth].synchronize()
        res = pipe.slots[f % depth].results()
        got.append([res[s] for s in check])

    for f in range(F):
        if f >= depth:
            collect(f - depth)
        pipe.submit(fid[f], raw[f], [c[f] for c in heads[0]], [c[f] for c in heads[1]], [c[f] for c in heads[2]],
                    done_event=done[f % depth])
    for f in range(max(F - depth, 0), F):
        collect(f)
    pipe.join()
    torch.cuda.synchronize()
    assert eng.status().max() == 0
    rows = evicted = 0
    max_id = -1
    for k, s in enumerate(check):
        o = m.StereoTrackOracle()
        alive_prev = set()
        for f in range(F):
            ref = o.step(f, host_raw[f, k], [c[f, k] for c in host_heads[0]], [c[f, k] for c in host_heads[1]],
                         [c[f, k] for c in host_heads[2]])
            _rows_equal(got[f][k], ref, f'frame {f} seq {s}')
            rows += len(ref['instances_id'])
            alive = set(o.tracker.export_state()['ids'].tolist())
            evicted += len(alive_prev - alive)
            alive_prev = alive
            max_id = max([max_id] + ref['instances_id'].tolist())
        _state_equal(eng.export_state(s), o.tracker.export_state())
    # the run must have exercised eviction and id growth, not just steady tracking
    assert rows > 8000 and evicted > 20 and max_id > 60, (rows, evicted, max_id)


def test_c1_full_length_300_frames_matches_oracle(mods):
    """BASELINE.json configs[0]: the default tracker on synthetic 640x640 sequences, all 300
    frames; every output row, and the whole tracker state (Kalman means / covariances, flags,
    counters, velocities) every 10th frame."""
    m, torch = mods, mods.torch
    cfg, S, F = m.CONFIGS['C1'], 2, 300
    fr = m.SyntheticStereo(cfg, S, F, seed=20240604).frames(0, F)
    raw = fr['raw'].cuda().contiguous()
    heads = [[c.cuda().contiguous() for c in fr[k]] for k in ('cls', 'bbox', 'obj')]
    eng = m.StereoTrackEngine(m.EngineConfig(n_seq=S))
    oracles = [m.StereoTrackOracle() for _ in range(S)]
    rows = 0
    for f in range(F):
        eng.frame_ids.fill_(f)
        eng.step(raw[f], [c[f] for c in heads[0]], [c[f] for c in heads[1]], [c[f] for c in heads[2]])
        res = eng.results()
        for s in range(S):
            ref = oracles[s].step(f, fr['raw'][f, s].numpy(), [c[f, s].numpy() for c in fr['cls']],
                                  [c[f, s].numpy() for c in fr['bbox']], [c[f, s].numpy() for c in fr['obj']])
            _rows_equal(res[s], ref, f'frame {f} seq {s}')
            rows += len(ref['instances_id'])
            if f % 10 == 0 or f == F - 1:
                _state_equal(eng.export_state(s), oracles[s].tracker.export_state())
    assert rows > 2000 and eng.status().max() == 0
