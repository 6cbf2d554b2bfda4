"""Synthetic AirDrone-shaped stereo sequences.

Produces, for S independent sequences, the two inputs of the hot path per
frame: the raw ``uint16`` disparity map (16 * pixels, 65535 = invalid; the file
format read at mmtrack/datasets/transforms/loading_disparity.py:77-86) and the
three YOLOX head maps (cls / bbox / obj logits at strides 8, 16, 32; the
tensors mmyolo's ``YOLOXHead.predict_by_feat`` receives at
mmtrack/models/detectors/yolo_detector_disparity_v1.py:121).

Written with torch ops so the same code fills host tensors for the parity tests
and device tensors for ``bench.py``; trajectories are closed-form in the frame
index so any frame range can be generated independently.
"""
from dataclasses import dataclass

import math
import torch

STRIDES = (8, 16, 32)


@dataclass
class SceneConfig:
    height: int = 720
    width: int = 1280
    k_min: int = 5
    k_max: int = 50
    invalid_prob: float = 0.05
    n_clutter: int = 40
    miss_rate: float = 0.05
    heavy_ego: bool = False
    fps: float = 10.0
    focal_disp: float = 640.0      # ocsort_disparity.py:36-37 (model constants)
    baseline: float = 0.25
    size_divisor: int = 32

    @property
    def padded_hw(self):
        d = self.size_divisor
        return (-(-self.height // d) * d, -(-self.width // d) * d)

    @property
    def level_hw(self):
        ph, pw = self.padded_hw
        return [(ph // s, pw // s) for s in STRIDES]

    @property
    def num_priors(self):
        return sum(h * w for h, w in self.level_hw)


CONFIGS = {
    # BASELINE.json configs[0..3] (C1..C4)
    'C1': SceneConfig(640, 640, 5, 10),
    'C2': SceneConfig(720, 1280, 5, 50),
    'C3': SceneConfig(1080, 1920, 20, 200, n_clutter=60),
    'C4': SceneConfig(720, 1280, 5, 50, invalid_prob=0.30, n_clutter=550, heavy_ego=True),
}


class SyntheticStereo:
    def __init__(self, cfg: SceneConfig, n_seq: int, n_frames: int, seed: int = 20240604,
                 device='cpu'):
        self.cfg, self.S, self.F = cfg, n_seq, n_frames
        self.device = torch.device(device)
        self.seed = seed
        g = torch.Generator(device='cpu').manual_seed(seed)
        S, K, H, W = n_seq, cfg.k_max, cfg.height, cfg.width
        f = W / 2.0
        U = lambda *shape: torch.rand(*shape, generator=g, dtype=torch.float64)
        N = lambda *shape: torch.randn(*shape, generator=g, dtype=torch.float64)
        n_obj = torch.randint(cfg.k_min, cfg.k_max + 1, (S,), generator=g)
        self.active = (torch.arange(K)[None, :] < n_obj[:, None])
        z0 = 3 + 77 * U(S, K)
        u0 = (0.1 + 0.8 * U(S, K)) * W
        v0 = (0.1 + 0.8 * U(S, K)) * H
        self.p0 = torch.stack([(u0 - W / 2) * z0 / f, (v0 - H / 2) * z0 / f, z0], -1)
        self.vel = 2.0 * N(S, K, 3)
        self.size = 0.3 + 0.9 * U(S, K)
        self.aspect = 0.8 + 0.4 * U(S, K)
        # ego translation random walk (only used when heavy_ego)
        self.ego_t = torch.cumsum(6.0 * N(S, n_frames, 2), dim=1)
        # per-sequence static background noise in raw (1/16 px) units
        self.bg_noise = N(S, H, W).to(torch.float32)

    # ---- geometry -------------------------------------------------------
    def gt_boxes(self, f0, f1):
        """Returns (boxes [F,S,K,4] xyxy float64, visible [F,S,K], z [F,S,K])."""
        cfg = self.cfg
        H, W = cfg.height, cfg.width
        f = W / 2.0
        t = torch.arange(f0, f1, dtype=torch.float64) / cfg.fps
        p = self.p0[None] + self.vel[None] * t[:, None, None, None]
        z = p[..., 2].clamp(3.0, 80.0)
        u = W / 2 + f * p[..., 0] / z
        v = H / 2 + f * p[..., 1] / z
        side = (f * self.size[None] / z).clamp(min=11.0)
        if cfg.heavy_ego:
            fr = torch.arange(f0, f1, dtype=torch.float64)
            s = 1 + 0.02 * torch.sin(fr / 7)
            a = math.radians(0.5) * torch.sin(fr / 11)
            ca, sa = torch.cos(a)[:, None, None], torch.sin(a)[:, None, None]
            du, dv = u - W / 2, v - H / 2
            tt = self.ego_t[:, f0:f1].permute(1, 0, 2)
            u = s[:, None, None] * (ca * du - sa * dv) + W / 2 + tt[..., 0:1]
            v = s[:, None, None] * (sa * du + ca * dv) + H / 2 + tt[..., 1:2]
            side = side * s[:, None, None]
        w = side * self.aspect[None]
        h = side
        boxes = torch.stack([u - w / 2, v - h / 2, u + w / 2, v + h / 2], -1)
        vis = self.active[None] & (boxes[..., 2] > 2) & (boxes[..., 3] > 2) & \
            (boxes[..., 0] < W - 2) & (boxes[..., 1] < H - 2)
        return boxes, vis, z

    # ---- frame synthesis ---------------------------------------------------
    def frames(self, f0, f1):
        """dict(raw [F,S,H,W] u16, cls/bbox/obj: per-level lists of
        [F,S,C,h,w] f32, gt_boxes, gt_visible) on ``self.device``."""
        # clusters of neighbouring objects write the same head-map cells: indexed
        # assignment with duplicate indices is only reproducible on CUDA in
        # deterministic mode
        prev = torch.are_deterministic_algorithms_enabled()
        torch.use_deterministic_algorithms(True, warn_only=True)
        try:
            return self._frames(f0, f1)
        finally:
            torch.use_deterministic_algorithms(prev)

    def _frames(self, f0, f1):
        cfg, dev = self.cfg, self.device
        S, K, H, W = self.S, cfg.k_max, cfg.height, cfg.width
        F = f1 - f0
        g = torch.Generator(device=dev).manual_seed(self.seed * 7919 + f0 * 31 + 1)
        boxes, vis, z = self.gt_boxes(f0, f1)
        boxes_d, vis_d, z_d = boxes.to(dev), vis.to(dev), z.to(dev)

        # ---- disparity -----------------------------------------------------
        rows = torch.arange(H, device=dev, dtype=torch.float32)
        z_bg = 400.0 - (340.0 / max(H - 1, 1)) * rows          # far at the top
        bg = 16.0 * cfg.focal_disp * cfg.baseline / z_bg       # [H]
        noise = self.bg_noise.to(dev)                          # [S,H,W]
        # per-frame variation: cyclic shift of the static noise field
        raw = torch.stack([bg[None, :, None] + torch.roll(
            noise, shifts=(3 * (f0 + i), 5 * (f0 + i)), dims=(1, 2)) for i in range(F)])
        raw = raw.round().clamp(0, 65534)
        obj_raw = (16.0 * cfg.focal_disp * cfg.baseline / z_d).round().to(torch.float32)
        ys = torch.arange(H, device=dev, dtype=torch.float32)
        xs = torch.arange(W, device=dev, dtype=torch.float32)
        x1 = boxes_d[..., 0].floor().to(torch.float32)
        y1 = boxes_d[..., 1].floor().to(torch.float32)
        x2 = boxes_d[..., 2].ceil().to(torch.float32)
        y2 = boxes_d[..., 3].ceil().to(torch.float32)
        for k in range(K):
            if not bool(vis[:, :, k].any()):
                continue
            in_y = (ys[None, None] >= y1[:, :, k, None]) & (ys[None, None] < y2[:, :, k, None])
            in_x = (xs[None, None] >= x1[:, :, k, None]) & (xs[None, None] < x2[:, :, k, None])
            val = torch.where(vis_d[:, :, k], obj_raw[:, :, k], torch.zeros((), device=dev))
            layer = (in_y[..., :, None] & in_x[..., None, :]).to(torch.float32) * val[..., None, None]
            raw = torch.maximum(raw, layer)
        bh, bw = -(-H // 8), -(-W // 8)
        inv = torch.rand(F, S, bh, bw, generator=g, device=dev) < cfg.invalid_prob
        inv = inv.repeat_interleave(8, 2).repeat_interleave(8, 3)[:, :, :H, :W]
        raw = torch.where(inv, torch.full((), 65535.0, device=dev), raw)
        # uint16 has no arithmetic in torch: stage through int32
        raw_u16 = raw.to(torch.int32).to(torch.uint16)

        # ---- head maps -------------------------------------------------------
        lv_hw = cfg.level_hw
        cls_l, box_l, obj_l = [], [], []
        for (h, w) in lv_hw:
            cls_l.append(torch.randn(F, S, 1, h, w, generator=g, device=dev) - 2.0)
            obj_l.append(torch.randn(F, S, 1, h, w, generator=g, device=dev) - 7.0)
            box_l.append(torch.randn(F, S, 4, h, w, generator=g, device=dev))
        bw_ = (boxes_d[..., 2] - boxes_d[..., 0]).to(torch.float32)
        bh_ = (boxes_d[..., 3] - boxes_d[..., 1]).to(torch.float32)
        cx = ((boxes_d[..., 0] + boxes_d[..., 2]) / 2).to(torch.float32)
        cy = ((boxes_d[..., 1] + boxes_d[..., 3]) / 2).to(torch.float32)
        side = torch.maximum(bw_, bh_)
        level = (side > 48).to(torch.int64) + (side > 96).to(torch.int64)
        detected = vis_d & (torch.rand(F, S, K, generator=g, device=dev) >= cfg.miss_rate)
        fi, si, ki = torch.nonzero(detected, as_tuple=True)
        for li, ((h, w), s) in enumerate(zip(lv_hw, STRIDES)):
            m = level[fi, si, ki] == li
            if not bool(m.any()):
                continue
            f_, s_, k_ = fi[m], si[m], ki[m]
            gx = (cx[f_, s_, k_] / s).floor().to(torch.int64)
            gy = (cy[f_, s_, k_] / s).floor().to(torch.int64)
            off = torch.tensor([-1, 0, 1], device=dev)
            px = (gx[:, None, None] + off[None, None, :]).expand(-1, 3, 3).reshape(-1)
            py = (gy[:, None, None] + off[None, :, None]).expand(-1, 3, 3).reshape(-1)
            rep = lambda t: t[:, None].expand(-1, 9).reshape(-1)
            f9, s9, k9 = rep(f_), rep(s_), rep(k_)
            ok = (px >= 0) & (px < w) & (py >= 0) & (py < h)
            px, py, f9, s9, k9 = px[ok], py[ok], f9[ok], s9[ok], k9[ok]
            n = px.numel()
            nz = lambda: torch.randn(n, generator=g, device=dev)
            cls_l[li][f9, s9, 0, py, px] = 3.0 + nz()
            obj_l[li][f9, s9, 0, py, px] = 3.0 + nz()
            box_l[li][f9, s9, 0, py, px] = (cx[f9, s9, k9] - px * s) / s + 0.05 * nz()
            box_l[li][f9, s9, 1, py, px] = (cy[f9, s9, k9] - py * s) / s + 0.05 * nz()
            box_l[li][f9, s9, 2, py, px] = torch.log(bw_[f9, s9, k9] / s) + 0.05 * nz()
            box_l[li][f9, s9, 3, py, px] = torch.log(bh_[f9, s9, k9] / s) + 0.05 * nz()
        if cfg.n_clutter > 0:
            C = cfg.n_clutter
            (h, w), s = lv_hw[0], STRIDES[0]
            px = torch.randint(0, w, (F, S, C), generator=g, device=dev).reshape(-1)
            py = torch.randint(0, h, (F, S, C), generator=g, device=dev).reshape(-1)
            f9 = torch.arange(F, device=dev)[:, None, None].expand(F, S, C).reshape(-1)
            s9 = torch.arange(S, device=dev)[None, :, None].expand(F, S, C).reshape(-1)
            n = px.numel()
            nz = lambda: torch.randn(n, generator=g, device=dev)
            cls_l[0][f9, s9, 0, py, px] = 0.5 * nz()
            obj_l[0][f9, s9, 0, py, px] = 0.5 * nz()
            box_l[0][f9, s9, 2, py, px] = 0.8 + 0.4 * nz()
            box_l[0][f9, s9, 3, py, px] = 0.8 + 0.4 * nz()
        return dict(raw=raw_u16, cls=cls_l, bbox=box_l, obj=obj_l,
                    gt_boxes=boxes_d, gt_visible=vis_d)
