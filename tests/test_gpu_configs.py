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
