"""Lockstep engine: S independent sequences advance one frame per call.

Host side of the C ABI (include/stereotrack_b200.h).  PyTorch is used for
device memory and streams only; every arithmetic step of the path runs in the
CUDA library.  The per-frame pipeline mirrors OCSORT_Disparity.predict
(mmtrack/models/mot/ocsort_disparity.py:79-109) after the detector backbone:

  raw disparity --K1--> depth map
  head maps     --K2--> detections (decode, threshold, NMS)
  detections    --K3--> per-box depth, scale, inflated boxes      (:82-83)
  inflated dets --K4/K5--> tracker.track                          (:88-93)
  tracks        --scale_boxes(1/scale)--> deflated boxes          (:96-97)
  deflated      --K3--> output depth                              (:100-104)
"""
import ctypes as C
from dataclasses import dataclass, field

import torch

from . import _lib

# the shipped tracker config: configs/stereo_tracking/ocsort/
# yolox_s_mmyolo_mot_airdrone_disp.py:48-58 (ctor defaults differ, see _lib.StConfig)
AIRDRONE_TRACKER = dict(obj_score_thr=0.3, init_track_thr=0.7, weight_iou_with_det_scores=False,
                        match_iou_thr=0.1, num_tentatives=3, vel_consist_weight=0.2, vel_delta_t=3,
                        num_frames_retain=30)


@dataclass
class EngineConfig:
    n_seq: int = 1
    max_tracks: int = 256
    max_dets: int = 256
    max_candidates: int = 2048
    tracker: dict = field(default_factory=lambda: dict(AIRDRONE_TRACKER))
    baseline: float = 0.25
    focal_length: float = 640.0
    score_thr: float = 0.01
    nms_iou_thr: float = 0.5
    strides: tuple = (8, 16, 32)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


class FrameBuffers:
    """One set of per-frame device buffers (detections, tracker outputs, K3 scratch,
    depth map): what one st_frame_io points at besides the inputs."""

    def __init__(self, S, D, dev):
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        i64 = dict(dtype=torch.int64, device=dev)
        # detector output (K2) and per-detection depth / scale (K3)
        self.det_boxes = torch.zeros(S, D, 4, **f32)
        self.det_scores = torch.zeros(S, D, **f32)
        self.det_labels = torch.zeros(S, D, **i64)
        self.det_prior = torch.zeros(S, D, **i32)
        self.det_keep = torch.zeros(S, D, **i32)
        self.det_counts = torch.zeros(S, **i32)
        self.det_depth = torch.zeros(S, D, **f32)
        self.det_scales = torch.zeros(S, D, **f32)
        self.det_inflated = torch.zeros(S, D, 4, **f32)
        # track outputs.  The rows a caller reads back (pred_track_instances fields + counts) are
        # views of ONE allocation, so a frame's results cross PCIe as a single copy (out_block).
        self.trk_boxes = torch.zeros(S, D, 4, **f32)       # inflated (tracker space)
        fields = [('out_boxes', (S, D, 4), torch.float32), ('out_ids', (S, D), torch.int64),
                  ('out_scores', (S, D), torch.float32), ('out_labels', (S, D), torch.int64),
                  ('out_scales', (S, D), torch.float32), ('out_depth', (S, D), torch.float32),
                  ('out_counts', (S,), torch.int32)]
        offs, total = [], 0
        for _, shape, dt in fields:
            offs.append(total)
            n = 1
            for v in shape:
                n *= v
            total += -(-n * torch.empty(0, dtype=dt).element_size() // 256) * 256
        self.out_block = torch.zeros(total, dtype=torch.uint8, device=dev)
        self.out_layout = []
        for (name, shape, dt), o in zip(fields, offs):
            n = 1
            for v in shape:
                n *= v
            nbytes = n * torch.empty(0, dtype=dt).element_size()
            setattr(self, name, self.out_block[o:o + nbytes].view(dt).view(*shape))
            self.out_layout.append((name, o, shape, dt))
        self.trk_depth = torch.zeros(S, D, **f32)          # depth carried through the tracker
        self.out_scale2 = torch.zeros(S, D, **f32)
        self.out_src = torch.zeros(S, D, **i32)            # detection row of every track row
        self.frame_ids = torch.zeros(S, **i32)
        self.k3_workspace = torch.zeros(8 + 6 * S * D, **i32)
        self.depth_map = None

    def outputs(self):
        return [self.out_boxes, self.out_ids, self.out_scores, self.out_labels, self.out_scales,
                self.out_depth, self.out_counts]

    @staticmethod
    def unpack(block, layout):
        """The output tensors as views of a (host or device) copy of ``out_block``."""
        out = {}
        for name, o, shape, dt in layout:
            n = 1
            for v in shape:
                n *= v
            out[name] = block[o:o + n * torch.empty(0, dtype=dt).element_size()].view(dt).view(*shape)
        return out

    def results(self, host_block=None):
        """Host copies of the outputs, one dict per sequence (``host_block``: a host copy of
        ``out_block`` that was already transferred)."""
        blk = self.out_block.cpu() if host_block is None else host_block
        v = self.unpack(blk, self.out_layout)
        n = v['out_counts'].tolist()
        pack = [v[k] for k in ('out_boxes', 'out_ids', 'out_scores', 'out_labels', 'out_scales', 'out_depth')]
        out = []
        for s, m in enumerate(n):
            out.append(dict(bboxes=pack[0][s, :m].numpy(), instances_id=pack[1][s, :m].numpy(),
                            scores=pack[2][s, :m].numpy(), labels=pack[3][s, :m].numpy(),
                            scales=pack[4][s, :m].numpy(), depth=pack[5][s, :m].numpy()))
        return out


class FrameInputs:
    """One lockstep frame's inputs - frame ids [S] i32, raw disparity [S,H,W] u16, head maps per
    level cls [S,C,h,w] / bbox [S,4,h,w] / obj [S,1,h,w] f32 - as views of ONE allocation
    (``block``, uint8; fields 256-byte aligned), on the host (optionally pinned) or a device: a
    frame crosses PCIe as a single ``cudaMemcpyAsync`` (``dst.block.copy_(src.block)``) and lands
    at fixed device addresses, so the frame's graph needs no pointer update."""

    def __init__(self, S, H, W, level_shapes, num_classes=1, device='cpu', pin=False):
        fields = [('frame_ids', (S,), torch.int32), ('raw', (S, H, W), torch.uint16)]
        for l, (h, w) in enumerate(level_shapes):
            fields += [(f'cls{l}', (S, num_classes, h, w), torch.float32), (f'bbox{l}', (S, 4, h, w), torch.float32),
                       (f'obj{l}', (S, 1, h, w), torch.float32)]
        offs, total = [], 0
        for _, shape, dt in fields:
            offs.append(total)
            n = 1
            for v in shape:
                n *= v
            total += -(-n * torch.empty(0, dtype=dt).element_size() // 256) * 256
        self.block = torch.zeros(total, dtype=torch.uint8, device=device)
        if pin:
            self.block = self.block.pin_memory()
        v = {}
        for (name, shape, dt), o in zip(fields, offs):
            n = 1
            for k in shape:
                n *= k
            v[name] = self.block[o:o + n * torch.empty(0, dtype=dt).element_size()].view(dt).view(*shape)
        L = len(level_shapes)
        self.frame_ids, self.raw = v['frame_ids'], v['raw']
        self.cls = [v[f'cls{l}'] for l in range(L)]
        self.bbox = [v[f'bbox{l}'] for l in range(L)]
        self.obj = [v[f'obj{l}'] for l in range(L)]
        self.nbytes = total

    def fill(self, frame_id, raw, cls_l, bbox_l, obj_l):
        self.frame_ids.fill_(int(frame_id))
        self.raw.copy_(raw)
        for dst, src in zip(self.cls + self.bbox + self.obj, list(cls_l) + list(bbox_l) + list(obj_l)):
            dst.copy_(src)
        return self


class StereoTrackEngine:
    def __init__(self, cfg: EngineConfig = None, device='cuda:0', **kw):
        cfg = cfg or EngineConfig(**kw)
        if not torch.cuda.is_available():
            raise _lib.StereoTrackError('StereoTrackEngine needs a CUDA device (no CPU fallback)')
        self.cfg = cfg
        self.device = torch.device(device)
        self.lib = _lib.lib()
        sc = _lib.StConfig()
        self.lib.st_default_config(C.byref(sc))
        sc.max_sequences, sc.max_tracks, sc.max_dets = cfg.n_seq, cfg.max_tracks, cfg.max_dets
        sc.max_candidates = cfg.max_candidates
        t = cfg.tracker
        # constructor defaults of OCSORTTracker_Disparity when a key is absent
        sc.obj_score_thr = t.get('obj_score_thr', 0.3)
        sc.init_track_thr = t.get('init_track_thr', 0.7)
        sc.weight_iou_with_det_scores = int(bool(t.get('weight_iou_with_det_scores', True)))
        sc.match_iou_thr = float(t.get('match_iou_thr', 0.3))
        sc.num_tentatives = t.get('num_tentatives', 3)
        sc.vel_consist_weight = t.get('vel_consist_weight', 0.2)
        sc.vel_delta_t = t.get('vel_delta_t', 3)
        sc.num_frames_retain = t.get('num_frames_retain', 10)
        sc.baseline, sc.focal_length = cfg.baseline, cfg.focal_length
        sc.score_thr, sc.nms_iou_thr = cfg.score_thr, cfg.nms_iou_thr
        self._sc = sc
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.st_create(C.byref(sc), C.byref(self._h)), 'st_create')
        # the engine's own buffer set; its tensors are also attributes of the engine
        self.buffers = FrameBuffers(cfg.n_seq, cfg.max_dets, self.device)
        self.__dict__.update(vars(self.buffers))

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self.lib.st_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t, dtype):
        assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), (t.device, t.dtype, t.is_contiguous())
        return t

    # ---- K1 ----------------------------------------------------------------------
    def disp2depth(self, raw, resized_hw=None, size_divisor=32, want_disp3=False, want_mask=False, out=None):
        """raw [S,H,W] uint16 -> depth [S,Hp,Wp] (+disp3 [S,3,Hp,Wp], mask [S,Hp,Wp])."""
        self._dev(raw, torch.uint16)
        S, H, W = raw.shape
        Hr, Wr = resized_hw or (H, W)
        Hp, Wp = -(-Hr // size_divisor) * size_divisor, -(-Wr // size_divisor) * size_divisor
        depth = out if out is not None else torch.empty(S, Hp, Wp, dtype=torch.float32, device=raw.device)
        disp3 = torch.empty(S, 3, Hp, Wp, dtype=torch.float32, device=raw.device) if want_disp3 else None
        mask = torch.empty(S, Hp, Wp, dtype=torch.uint8, device=raw.device) if want_mask else None
        _lib.check(self.lib.st_disp2depth(_ptr(raw), S, H, W, Hr, Wr, Hp, Wp, _ptr(depth), _ptr(disp3),
                                          _ptr(mask), self.cfg.baseline, self.cfg.focal_length,
                                          self._stream()), 'st_disp2depth')
        return depth, disp3, mask

    def pack_image(self, img, size_divisor=32, pad_val=(114, 114, 114), mean=None, std=None, swap_rb=False,
                   out=None, resized_hw=None):
        """K1b: u8 frames [S,H,W,3] (device) -> the detector's float32 input [S,3,H',W'] (optional
        bilinear resize to ``resized_hw`` as cv2.resize INTER_LINEAR does it, pad with ``pad_val`` to a
        multiple of ``size_divisor``, HWC->CHW, optional channel swap and (img - mean) / std; see
        st_pack_image / st_resize_pack_image)."""
        import ctypes
        S, H, W, ch = img.shape
        if ch != 3 or img.dtype != torch.uint8:
            raise ValueError('img must be [S,H,W,3] uint8')
        if (mean is None) != (std is None):
            raise AssertionError('mean and std should be both None or tuple')      # data_preprocessor.py:65-66
        Hr, Wr = resized_hw or (H, W)
        Hp = -(-Hr // size_divisor) * size_divisor if size_divisor > 1 else Hr
        Wp = -(-Wr // size_divisor) * size_divisor if size_divisor > 1 else Wr
        out = out if out is not None else torch.empty(S, 3, Hp, Wp, dtype=torch.float32, device=img.device)
        pad = (ctypes.c_uint8 * 3)(*[int(v) for v in pad_val])
        m = (ctypes.c_float * 3)(*mean) if mean is not None else None
        sd = (ctypes.c_float * 3)(*std) if std is not None else None
        _lib.check(self.lib.st_resize_pack_image(_ptr(img.contiguous()), S, H, W, Hr, Wr, Hp, Wp, pad, m, sd,
                                                 int(bool(swap_rb)), _ptr(out), self._stream()), 'st_resize_pack_image')
        return out

    def disp_to_depth(self, disp, out=None):
        """float disparity (any shape, contiguous) -> depth, ocsort_disparity.py:132-134."""
        self._dev(disp, torch.float32)
        out = out if out is not None else torch.empty_like(disp)
        _lib.check(self.lib.st_disp_f32_to_depth(_ptr(disp), disp.numel(), _ptr(out), self.cfg.baseline,
                                                 self.cfg.focal_length, self._stream()), 'st_disp_f32_to_depth')
        return out

    # ---- K2 ----------------------------------------------------------------------
    def decode_nms(self, cls_l, bbox_l, obj_l, ori_hw, scale_factor=(1.0, 1.0)):
        """Per-level head maps [S,C,h,w] -> fills det_* buffers; returns views."""
        L = len(cls_l)
        S = cls_l[0].shape[0]
        nc = cls_l[0].shape[1]
        for t in list(cls_l) + list(bbox_l) + list(obj_l):
            self._dev(t, torch.float32)
        arr = lambda ts: (C.c_void_p * L)(*[t.data_ptr() for t in ts])
        ints = lambda vs: (C.c_int32 * L)(*vs)
        _lib.check(self.lib.st_decode_nms(
            self._h, arr(cls_l), arr(bbox_l), arr(obj_l), L, ints([t.shape[2] for t in cls_l]),
            ints([t.shape[3] for t in cls_l]), ints(self.cfg.strides[:L]), S, nc,
            float(scale_factor[0]), float(scale_factor[1]), int(ori_hw[0]), int(ori_hw[1]),
            self.cfg.max_dets, _ptr(self.det_boxes), _ptr(self.det_scores), _ptr(self.det_labels),
            _ptr(self.det_prior), _ptr(self.det_keep), _ptr(self.det_counts), self._stream()),
            'st_decode_nms')
        return self.det_boxes, self.det_scores, self.det_labels, self.det_counts

    # ---- K3 ----------------------------------------------------------------------
    DEPTH_MODES = dict(window=0, truncated_mean=1, mean=2, median=3, center=4)

    def box_depth(self, depth, boxes, counts, d_out=None, scale_out=None, inflated_out=None, mode='window'):
        """``mode``: 'window' = extract_depth (ocsort_disparity.py:136-175); 'truncated_mean',
        'mean', 'median', 'center' = the ablation reducers of depth_extraction_comparison.py
        (scale = clip(d*d/400, 1, 3)).  For 'center', ``self.depth_oob`` [S] int32 counts the boxes
        whose centre pixel lies outside the map (the reference raises IndexError there)."""
        if mode not in self.DEPTH_MODES:
            raise ValueError(f'unknown depth reducer {mode!r}')
        S, Hp, Wp = depth.shape
        cap = boxes.shape[1]
        dev = depth.device
        d_out = d_out if d_out is not None else torch.empty(S, cap, dtype=torch.float32, device=dev)
        scale_out = scale_out if scale_out is not None else torch.empty(S, cap, dtype=torch.float32, device=dev)
        ws = self.k3_workspace
        if ws.numel() < 8 + 6 * S * cap or ws.device != dev:
            ws = self.k3_workspace = torch.zeros(8 + 6 * S * cap, dtype=torch.int32, device=dev)
        if mode == 'window':
            _lib.check(self.lib.st_box_depth(_ptr(self._dev(depth, torch.float32)), S, Hp, Wp,
                                             _ptr(self._dev(boxes, torch.float32)),
                                             _ptr(self._dev(counts, torch.int32)), cap, _ptr(d_out),
                                             _ptr(scale_out), _ptr(inflated_out), _ptr(ws), self._stream()),
                       'st_box_depth')
        else:
            self.depth_oob = torch.empty(S, dtype=torch.int32, device=dev)
            _lib.check(self.lib.st_box_depth_mode(_ptr(self._dev(depth, torch.float32)), S, Hp, Wp,
                                                  _ptr(self._dev(boxes, torch.float32)),
                                                  _ptr(self._dev(counts, torch.int32)), cap,
                                                  self.DEPTH_MODES[mode], _ptr(d_out), _ptr(scale_out),
                                                  _ptr(inflated_out), _ptr(self.depth_oob), _ptr(ws),
                                                  self._stream()),
                       'st_box_depth_mode')
        return d_out, scale_out, inflated_out

    def box_depth_raw(self, raw, padded_hw, boxes, counts, d_out=None, scale_out=None, inflated_out=None,
                      prev_boxes=None, prev_index=None, prev_d=None):
        """extract_depth (ocsort_disparity.py:136-175) straight from the raw u16 frames [S,H,W]:
        bit-identical to ``box_depth`` on the map ``disp2depth(raw)`` writes (resize factor 1),
        without that map (st_box_depth_raw).  ``prev_*``: reuse of an earlier call's values as
        in st_box_depth_reuse."""
        S, H, W = raw.shape
        Hp, Wp = padded_hw
        cap = boxes.shape[1]
        dev = raw.device
        d_out = d_out if d_out is not None else torch.empty(S, cap, dtype=torch.float32, device=dev)
        scale_out = scale_out if scale_out is not None else torch.empty(S, cap, dtype=torch.float32, device=dev)
        ws = self.k3_workspace
        if ws.numel() < 8 + 6 * S * cap or ws.device != dev:
            ws = self.k3_workspace = torch.zeros(8 + 6 * S * cap, dtype=torch.int32, device=dev)
        _lib.check(self.lib.st_box_depth_raw(_ptr(self._dev(raw, torch.uint16)), S, H, W, Hp, Wp, self.cfg.baseline,
                                             self.cfg.focal_length, _ptr(self._dev(boxes, torch.float32)),
                                             _ptr(self._dev(counts, torch.int32)), cap, _ptr(d_out), _ptr(scale_out),
                                             _ptr(inflated_out), _ptr(ws), _ptr(prev_boxes), _ptr(prev_index),
                                             _ptr(prev_d), self._stream()), 'st_box_depth_raw')
        return d_out, scale_out, inflated_out

    def scale_boxes(self, boxes, scales, counts, invert, out=None):
        S, cap = scales.shape
        out = out if out is not None else torch.empty_like(boxes)
        _lib.check(self.lib.st_scale_boxes(_ptr(boxes), _ptr(scales), _ptr(counts), S, cap, int(invert),
                                           _ptr(out), self._stream()), 'st_scale_boxes')
        return out

    # ---- K4/K5 -------------------------------------------------------------------
    def track_step(self, frame_ids, boxes, scores, labels, scales, depth, counts, warp=None):
        """tracker.track for S sequences; inputs padded [S,cap,...]; fills trk_*/out_* buffers.
        ``warp``: [S,6] float32 camera-motion affines (NaN row = none) when cmc is on."""
        if warp is not None:
            self._dev(warp, torch.float32)
            assert warp.shape == (scores.shape[0], 6)
        S, cap = scores.shape
        assert cap == self.cfg.max_dets, 'detection buffers must be padded to max_dets rows'
        _lib.check(self.lib.st_track_step(
            self._h, S, _ptr(self._dev(frame_ids, torch.int32)), _ptr(self._dev(boxes, torch.float32)),
            _ptr(self._dev(scores, torch.float32)), _ptr(self._dev(labels, torch.int64)),
            _ptr(self._dev(scales, torch.float32)), _ptr(self._dev(depth, torch.float32)),
            _ptr(self._dev(counts, torch.int32)), cap, _ptr(self.trk_boxes), _ptr(self.out_scores),
            _ptr(self.out_labels), _ptr(self.out_scales), _ptr(self.trk_depth), _ptr(self.out_ids),
            _ptr(self.out_counts), _ptr(self.out_src), _ptr(warp), self._stream()), 'st_track_step')
        return dict(bboxes=self.trk_boxes, scores=self.out_scores, labels=self.out_labels,
                    scales=self.out_scales, depth=self.trk_depth, instances_id=self.out_ids,
                    counts=self.out_counts)

    # ---- whole frame ---------------------------------------------------------------
    def step_from_dets(self, depth_map, frame_ids=None, warp=None):
        """ocsort_disparity.py:82-109 given det_* buffers (detector output) and a depth map."""
        fid = self.frame_ids if frame_ids is None else frame_ids
        self.box_depth(depth_map, self.det_boxes, self.det_counts, self.det_depth, self.det_scales,
                       self.det_inflated)
        self.track_step(fid, self.det_inflated, self.det_scores, self.det_labels, self.det_scales,
                        self.det_depth, self.det_counts, warp=warp)
        self.scale_boxes(self.trk_boxes, self.out_scales, self.out_counts, True, self.out_boxes)
        self.box_depth(depth_map, self.out_boxes, self.out_counts, self.out_depth, self.out_scale2, None)

    def step(self, raw, cls_l, bbox_l, obj_l, frame_ids=None, resized_hw=None):
        """One lockstep frame from raw inputs.  Results are left in the out_* buffers."""
        S, H, W = raw.shape
        Hr, Wr = resized_hw or (H, W)
        Hp, Wp = -(-Hr // 32) * 32, -(-Wr // 32) * 32
        if self.depth_map is None or tuple(self.depth_map.shape) != (S, Hp, Wp):
            self.depth_map = self.buffers.depth_map = torch.empty(S, Hp, Wp, dtype=torch.float32, device=self.device)
        self.disp2depth(raw, resized_hw, out=self.depth_map)
        self.decode_nms(cls_l, bbox_l, obj_l, (H, W))
        self.step_from_dets(self.depth_map, frame_ids)

    def make_frame_io(self, raw_shape, level_shapes, num_classes=1, resized_hw=None, scale_factor=(1.0, 1.0),
                      bufs=None, depth_source='map'):
        """Prepare a reusable st_frame_io for inputs of the given shapes: raw
        [S,H,W] and per-level head maps [S,C,h,w].  Pointers to the per-step inputs
        are filled by ``frame_step``; everything else is bound once to the buffer
        set ``bufs`` (default: the engine's own).  ``depth_source``: 'map' = K1 writes the
        float32 depth map and K3 gathers from it; 'raw' = K3 gathers from the raw frame and no
        map is written (resize factor 1 only); 'raw+map' = K3 from the raw frame, map written too."""
        if depth_source not in ('map', 'raw', 'raw+map'):
            raise ValueError(f'unknown depth_source {depth_source!r}')
        S, H, W = raw_shape
        Hr, Wr = resized_hw or (H, W)
        if depth_source != 'map' and (Hr, Wr) != (H, W):
            raise ValueError("depth_source='raw' needs resize factor 1 (st_box_depth_raw)")
        Hp, Wp = -(-Hr // 32) * 32, -(-Wr // 32) * 32
        own = bufs is None
        bufs = self if own else bufs
        if depth_source != 'raw' and (bufs.depth_map is None or tuple(bufs.depth_map.shape) != (S, Hp, Wp)):
            bufs.depth_map = torch.empty(S, Hp, Wp, dtype=torch.float32, device=self.device)
            if own:
                self.buffers.depth_map = bufs.depth_map
        L = len(level_shapes)
        io = _lib.StFrameIO()
        keep = dict(cls=(C.c_void_p * L)(), bbox=(C.c_void_p * L)(), obj=(C.c_void_p * L)(),
                    lh=(C.c_int32 * L)(*[s[0] for s in level_shapes]),
                    lw=(C.c_int32 * L)(*[s[1] for s in level_shapes]),
                    st=(C.c_int32 * L)(*self.cfg.strides[:L]))
        io.frame_ids = bufs.frame_ids.data_ptr()
        io.H, io.W, io.Hr, io.Wr, io.Hp, io.Wp = H, W, Hr, Wr, Hp, Wp
        io.cls, io.bbox, io.obj = (C.cast(keep[k], C.c_void_p) for k in ('cls', 'bbox', 'obj'))
        io.level_h, io.level_w, io.strides = (C.cast(keep[k], C.c_void_p) for k in ('lh', 'lw', 'st'))
        io.num_levels, io.num_classes, io.ori_h, io.ori_w = L, num_classes, H, W
        io.scale_w, io.scale_h = float(scale_factor[0]), float(scale_factor[1])
        io.depth_map = 0 if depth_source == 'raw' else bufs.depth_map.data_ptr()
        io.depth_from_raw = int(depth_source != 'map')
        for name in ('det_boxes', 'det_scores', 'det_labels', 'det_prior', 'det_keep', 'det_counts',
                     'det_depth', 'det_scales', 'det_inflated', 'trk_boxes', 'trk_depth', 'out_scale2',
                     'out_src', 'k3_workspace',
                     'out_boxes', 'out_scores', 'out_labels', 'out_scales', 'out_depth', 'out_ids',
                     'out_counts'):
            setattr(io, name, getattr(bufs, name).data_ptr())
        io._keep = keep
        return io

    def frame_step(self, io, raw, cls_l, bbox_l, obj_l):
        """One lockstep frame through st_frame_step (a single C call, KERNELS_PER_FRAME launches)."""
        io.raw = raw.data_ptr()
        for l in range(len(cls_l)):
            io._keep['cls'][l] = cls_l[l].data_ptr()
            io._keep['bbox'][l] = bbox_l[l].data_ptr()
            io._keep['obj'][l] = obj_l[l].data_ptr()
        _lib.check(self.lib.st_frame_step(self._h, raw.shape[0], C.byref(io), self._stream()), 'st_frame_step')

    def frame_step_ptrs(self, io, n_seq, frame_ids_ptr, raw_ptr, cls_ptrs, bbox_ptrs, obj_ptrs):
        """frame_step with raw device addresses (ints) prepared ahead of time: no tensor
        views or torch ops on the per-frame host path, one C call."""
        io.frame_ids = frame_ids_ptr
        io.raw = raw_ptr
        k = io._keep
        for l in range(len(cls_ptrs)):
            k['cls'][l] = cls_ptrs[l]
            k['bbox'][l] = bbox_ptrs[l]
            k['obj'][l] = obj_ptrs[l]
        _lib.check(self.lib.st_frame_step(self._h, n_seq, C.byref(io), self._stream()), 'st_frame_step')

    KERNELS_PER_FRAME = 13     # K1, K2 x2 (score, sort+NMS), K3 x4 (classify, warp, table, overflow), K4, deflate, K3 x4

    def results(self):
        """Host copies of the current outputs, one dict per sequence."""
        return self.buffers.results()

    # ---- state ----------------------------------------------------------------------
    def reset(self, seq_mask=None):
        m = None
        if seq_mask is not None:
            m = torch.as_tensor(seq_mask, dtype=torch.uint8, device=self.device).contiguous()
        _lib.check(self.lib.st_reset(self._h, _ptr(m), self._stream()), 'st_reset')
        if m is not None:
            torch.cuda.current_stream(self.device).synchronize()

    def live_counts(self):
        """Live tracks per sequence (len(tracker.tracks)); synchronises the current stream."""
        import numpy as np
        n = np.zeros(self.cfg.n_seq, np.int32)
        _lib.check(self.lib.st_read_live_counts(self._h, n.ctypes.data_as(C.c_void_p), self.cfg.n_seq,
                                                self._stream()), 'st_read_live_counts')
        return n

    def status(self):
        """Sticky ST_STATUS_* bits per sequence; waits for all work on the device first."""
        import numpy as np
        st = np.zeros(self.cfg.n_seq, np.int32)
        _lib.check(self.lib.st_read_status(self._h, st.ctypes.data_as(C.c_void_p), self.cfg.n_seq))
        return st

    def save_state(self):
        """Snapshot of the whole tracker state (all sequences) as a uint8 NumPy array
        (st_snapshot_save); join any frame pipeline first."""
        import numpy as np
        n = int(self.lib.st_snapshot_bytes(self._h))
        if n < 0:
            _lib.check(n, 'st_snapshot_bytes')
        buf = np.empty(n, np.uint8)
        _lib.check(self.lib.st_snapshot_save(self._h, buf.ctypes.data, n, self._stream()), 'st_snapshot_save')
        return buf

    def load_state(self, buf):
        """Restore a snapshot taken by ``save_state`` on an engine of the same capacities."""
        import numpy as np
        buf = np.ascontiguousarray(np.frombuffer(buf, np.uint8) if isinstance(buf, (bytes, bytearray)) else buf,
                                   dtype=np.uint8)
        _lib.check(self.lib.st_snapshot_load(self._h, buf.ctypes.data, buf.size, self._stream()), 'st_snapshot_load')

    def export_state(self, seq):
        import numpy as np
        T = self.cfg.max_tracks
        ids = np.zeros(T, np.int64); mean = np.zeros((T, 8)); cov = np.zeros((T, 64))
        vel = np.zeros((T, 2), np.float32); flags = np.zeros(T, np.uint8)
        hits = np.zeros(T, np.int32); last = np.zeros(T, np.int32); nt = np.zeros(1, np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        n = _lib.check(self.lib.st_export_state(self._h, seq, p(ids), p(mean), p(cov), p(vel), p(flags),
                                                p(hits), p(last), p(nt)), 'st_export_state')
        return dict(ids=ids[:n], mean=mean[:n], cov=cov[:n].reshape(n, 8, 8), velocity=vel[:n],
                    tentative=(flags[:n] & 1).astype(bool), tracked=(flags[:n] & 2).astype(bool),
                    hits=hits[:n], last_frame=last[:n], num_tracks=int(nt[0]))

    # ---- exposed primitives for parity tests -----------------------------------------
    def kalman_batch(self, op, mean, cov, z=None):
        n = mean.shape[0]
        _lib.check(self.lib.st_kalman_batch(op, n, _ptr(mean), _ptr(cov), _ptr(z), self._stream()))

    def debug_words(self):
        """[S,32] diagnostic words: 5 = assignments solved, 6 = solved by the exact
        fallback, 8..13 = SM cycles of the last track step per phase."""
        import numpy as np
        a = np.zeros((self.cfg.n_seq, 32), np.int32)
        _lib.check(self.lib.st_read_debug(self._h, a.ctypes.data_as(C.c_void_p), self.cfg.n_seq))
        return a

    def lapjv_batch(self, cost, cost_limit, mode=0):
        B, R, Cc = cost.shape
        x = torch.empty(B, R, dtype=torch.int32, device=cost.device)
        y = torch.empty(B, Cc, dtype=torch.int32, device=cost.device)
        _lib.check(self.lib.st_lapjv_batch(self._h, B, R, Cc, _ptr(self._dev(cost, torch.float32)),
                                           float(cost_limit), int(mode), _ptr(x), _ptr(y), self._stream()))
        return x, y


class FramePipeline:
    """Software pipeline over consecutive frames (st_pipeline_* of the C ABI): K1 and
    K2 of a frame run side by side, then K3 on its detections (the "pre" segment, on the ring
    slot's own stream), the tracker (one stream, frames in order) and the deflate + K3 on its
    tracks ("post"), so several consecutive frames are in different stages at once.  ``depth`` complete buffer
    sets are cycled; frame k uses set k % depth and its results stay readable until
    frame k + depth is submitted."""

    def __init__(self, engine, raw_shape, level_shapes, depth=5, num_classes=1, depth_source='map'):
        self.engine, self.depth, self.depth_source = engine, depth, depth_source
        S = raw_shape[0]
        self.n_seq = S
        self.slots = [engine.buffers] + [FrameBuffers(S, engine.cfg.max_dets, engine.device) for _ in range(depth - 1)]
        self.ios = [engine.make_frame_io(raw_shape, level_shapes, num_classes, bufs=None if i == 0 else b,
                                         depth_source=depth_source)
                    for i, b in enumerate(self.slots)]
        self._p = C.c_void_p()
        with torch.cuda.device(engine.device):
            _lib.check(engine.lib.st_pipeline_create(engine._h, depth, C.byref(self._p)), 'st_pipeline_create')
        self.submitted = 0

    def close(self):
        if self._p.value:
            self.engine.lib.st_pipeline_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit_ptrs(self, frame_ids_ptr, raw_ptr, cls_ptrs, bbox_ptrs, obj_ptrs, ready_event=None, done_event=None,
                    warp_ptr=0):
        """Queue one frame given raw device addresses; returns the slot it uses.
        ``warp_ptr``: address of the frame's [S,6] camera-motion affines, 0 = cmc off."""
        slot = self.submitted % self.depth
        io = self.ios[slot]
        io.frame_ids = frame_ids_ptr
        io.raw = raw_ptr
        io.warp = warp_ptr
        k = io._keep
        for l in range(len(cls_ptrs)):
            k['cls'][l] = cls_ptrs[l]
            k['bbox'][l] = bbox_ptrs[l]
            k['obj'][l] = obj_ptrs[l]
        def ev(e):
            if e is None:
                return C.c_void_p(0)
            if not e.cuda_event:          # torch creates the CUDA event lazily at its first record
                e.record(torch.cuda.current_stream(self.engine.device))
            return C.c_void_p(e.cuda_event)
        _lib.check(self.engine.lib.st_pipeline_submit(self._p, self.n_seq, C.byref(io), ev(ready_event),
                                                      ev(done_event)), 'st_pipeline_submit')
        self.submitted += 1
        return slot

    def submit(self, frame_ids, raw, cls_l, bbox_l, obj_l, ready_event=None, done_event=None, warp=None):
        """Queue one frame given device tensors (frame_ids [S] int32, raw [S,H,W] u16,
        heads [S,C,h,w] f32, optional warp [S,6] f32); they must stay alive and unchanged
        until the frame is done."""
        return self.submit_ptrs(frame_ids.data_ptr(), raw.data_ptr(), [t.data_ptr() for t in cls_l],
                                [t.data_ptr() for t in bbox_l], [t.data_ptr() for t in obj_l],
                                ready_event, done_event, 0 if warp is None else warp.data_ptr())

    def use_graphs(self, enable=True):
        """Replay frames as CUDA graphs (st_pipeline_use_graphs): from a slot's second use on."""
        _lib.check(self.engine.lib.st_pipeline_use_graphs(self._p, int(bool(enable))), 'st_pipeline_use_graphs')

    def partition(self):
        """(tracker SMs, other SMs) of the pipeline's SM partition, (0, 0) when not partitioned (st_pipeline_partition)."""
        a, b = C.c_int(0), C.c_int(0)
        _lib.check(self.engine.lib.st_pipeline_partition(self._p, C.byref(a), C.byref(b)), 'st_pipeline_partition')
        return a.value, b.value

    def stats(self):
        """dict(graph_frames=frames replayed as graphs, captures=stream captures so far)."""
        a, b = C.c_longlong(0), C.c_longlong(0)
        _lib.check(self.engine.lib.st_pipeline_stats(self._p, C.byref(a), C.byref(b)), 'st_pipeline_stats')
        return dict(graph_frames=int(a.value), captures=int(b.value))

    STAGES = ('k1_disp2depth', 'k2_decode_nms', 'k3_box_depth_dets', 'k4_track_step', 'deflate_k3_tracks')

    def trace(self, n_frames):
        """Time-stamp the stages of the next ``n_frames`` submitted frames (0 = off)."""
        _lib.check(self.engine.lib.st_pipeline_trace(self._p, int(n_frames)), 'st_pipeline_trace')

    def read_trace(self, max_frames=4096):
        """[frames, 5 stages, (begin, end)] milliseconds since ``trace`` (synchronises)."""
        import numpy as np
        a = np.zeros((max_frames, 5, 2), np.float32)
        n = _lib.check(self.engine.lib.st_pipeline_read_trace(self._p, a.ctypes.data_as(C.c_void_p), max_frames),
                       'st_pipeline_read_trace')
        return a[:n]

    def fork(self, stream=None):
        stream = stream or torch.cuda.current_stream(self.engine.device)
        _lib.check(self.engine.lib.st_pipeline_fork(self._p, C.c_void_p(stream.cuda_stream)), 'st_pipeline_fork')

    def join(self, stream=None):
        stream = stream or torch.cuda.current_stream(self.engine.device)
        _lib.check(self.engine.lib.st_pipeline_join(self._p, C.c_void_p(stream.cuda_stream)), 'st_pipeline_join')


class LockstepGroups:
    """S sequences split into G independent lockstep groups (one handle each), every
    group advancing through its own FramePipeline.  Sequences never interact (the
    reference shards whole videos across ranks, mmtrack/datasets/samplers/
    video_sampler.py:50-69), so no synchronisation between groups happens on the data
    path.  One group per GPU is the default: the frame pipeline already overlaps the
    latency-bound stages with the bandwidth-bound ones."""

    def __init__(self, n_seq, n_groups=1, device='cuda:0', depth=5, depth_source='map', **cfg_kw):
        self.depth_source = depth_source
        n_groups = max(1, min(n_groups, n_seq))
        base, extra = divmod(n_seq, n_groups)
        self.sizes = [base + (1 if g < extra else 0) for g in range(n_groups)]
        self.offsets = [sum(self.sizes[:g]) for g in range(n_groups)]
        self.device = torch.device(device)
        self.depth = depth
        self.engines = [StereoTrackEngine(EngineConfig(n_seq=sz, **cfg_kw), device=device) for sz in self.sizes]
        self.pipes = None

    def bind(self, raw_shape, level_shapes, num_classes=1):
        _, H, W = raw_shape
        self.pipes = [FramePipeline(e, (sz, H, W), level_shapes, self.depth, num_classes, self.depth_source)
                      for e, sz in zip(self.engines, self.sizes)]

    def plan(self, raw, cls_l, bbox_l, obj_l, first_frame_id=0):
        """Precompute the device addresses of every (frame, group) slice of resident
        inputs raw [F,S,H,W], heads [F,S,C,h,w] and of a frame-id table, so that
        ``run_plan`` issues a frame with one C call per group and no torch ops."""
        F = raw.shape[0]
        self._fid = [torch.arange(first_frame_id, first_frame_id + F, dtype=torch.int32, device=self.device)
                     [:, None].repeat(1, sz).contiguous() for sz in self.sizes]
        plan = []
        for f in range(F):
            row = []
            for g, (a, sz) in enumerate(zip(self.offsets, self.sizes)):
                row.append((self._fid[g][f].data_ptr(), raw[f, a:a + sz].data_ptr(),
                            [c[f, a:a + sz].data_ptr() for c in cls_l], [c[f, a:a + sz].data_ptr() for c in bbox_l],
                            [c[f, a:a + sz].data_ptr() for c in obj_l]))
            plan.append(row)
        self._plan = plan
        self._plan_keepalive = (raw, cls_l, bbox_l, obj_l)
        return plan

    def run_plan(self, f, ready_event=None, done_events=None):
        """Queue planned frame f on every group's pipeline; returns the slot used."""
        slot = 0
        for g, (pipe, args) in enumerate(zip(self.pipes, self._plan[f])):
            slot = pipe.submit_ptrs(*args, ready_event=ready_event,
                                    done_event=None if done_events is None else done_events[g])
        return slot

    def join(self, stream=None):
        for p in self.pipes:
            p.join(stream)

    def fork(self, stream=None):
        for p in self.pipes:
            p.fork(stream)

    def use_graphs(self, enable=True):
        for p in self.pipes:
            p.use_graphs(enable)

    def status(self):
        import numpy as np
        return np.concatenate([e.status() for e in self.engines])

    def results(self, slot=None):
        """Host copies of the outputs of the frame in ring slot ``slot`` (default: the
        last submitted frame); call after ``join`` + synchronisation."""
        out = []
        for p in self.pipes:
            k = (p.submitted - 1) % p.depth if slot is None else slot
            out.extend(p.slots[k].results())
        return out

    @property
    def kernels_per_frame(self):
        # without the depth map K1 is not launched
        per = StereoTrackEngine.KERNELS_PER_FRAME - (1 if self.depth_source == 'raw' else 0)
        return per * len(self.engines)
