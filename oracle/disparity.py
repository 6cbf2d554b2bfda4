"""Row A1: raw u16 disparity -> (mask, disparity px, depth m), resize, pad.

Follows
  mmtrack/datasets/transforms/loading_disparity.py:82-86,:129-134
  mmtrack/datasets/transforms/transforms_disparity.py:52-70,:92-110 (nearest
      resize through mmcv.imrescale == cv2.resize INTER_NEAREST)
  mmtrack/datasets/transforms/transforms_disparity.py:177-194,:215-232 (pad)
  mmtrack/models/mot/ocsort_disparity.py:132-134 (disp2depth)
"""
import numpy as np

INVALID_RAW = 65535


def load_postprocess(raw_u16):
    """loading_disparity.py:82-83 (mask) and :129-134 (_post_processing_v2).

    Returns (disp f32 [H,W] in pixels, mask u8 [H,W])."""
    raw = np.asarray(raw_u16)
    assert raw.dtype == np.uint16 and raw.ndim == 2
    mask = (raw < INVALID_RAW).astype(np.uint8)
    disp = raw.astype(np.float32)
    disp[disp == INVALID_RAW] = 0
    disp = disp / np.float32(16.0)
    return disp, mask


def rescale_size(h, w, scale):
    """mmcv.image.geometric.rescale_size for a (a, b) tuple scale, keep_ratio.

    transforms_disparity.py:55-60 passes results['scale'] = img_scale."""
    long_edge, short_edge = max(scale), min(scale)
    factor = min(long_edge / max(h, w), short_edge / min(h, w))
    return int(h * float(factor) + 0.5), int(w * float(factor) + 0.5)


def nearest_index(dst_n, src_n):
    """cv2.resize(INTER_NEAREST) source index of every destination index:
    min(floor(dst * (src/dst_n)), src-1), scale evaluated in float64."""
    inv = 1.0 / (dst_n / float(src_n))
    idx = np.floor(np.arange(dst_n, dtype=np.float64) * inv).astype(np.int64)
    return np.minimum(idx, src_n - 1)


def nearest_resize(img, new_h, new_w):
    """transforms_disparity.py:52-70.  Values are NOT multiplied by the resize
    factor (the reference does not either)."""
    h, w = img.shape[:2]
    if (new_h, new_w) == (h, w):
        return img.copy()
    return img[nearest_index(new_h, h)][:, nearest_index(new_w, w)]


def _linear_coeffs(dst_n, src_n):
    """cv2.resize(INTER_LINEAR) on 8-bit images, one axis: source index and the two 11-bit
    fixed-point weights of every destination index (OpenCV imgproc resize.cpp, resize():
    fx = (float)((dx + 0.5) * scale - 0.5), floor, weights saturate_cast<short>(w * 2048) with
    round-half-even).  ``clamp`` = the horizontal pass's border rule (index clamped, weight
    moved to the first tap); the vertical pass keeps the weights and clamps the ROW indices."""
    scale = 1.0 / (dst_n / float(src_n))
    f = ((np.arange(dst_n, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    i0 = np.floor(f).astype(np.int64)
    f = (f - i0.astype(np.float32)).astype(np.float32)
    return i0, f


def bilinear_resize_u8(img, new_h, new_w):
    """The image branch of Resize_Disparity (transforms_disparity.py:24-26 -> mmdet Resize._resize_img
    -> mmcv.imrescale(..., interpolation='bilinear', backend='cv2') == cv2.resize(INTER_LINEAR)) for an
    8-bit [H,W,C] image, restated bit-exactly: horizontal pass to 11-bit fixed point, vertical pass
    ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2 >> 2.  A downscale by exactly 2 in both axes
    is the 2x2 box average (OpenCV switches INTER_LINEAR to INTER_AREA there).  Pinned against
    cv2.resize by tests/golden/make_golden.py (ref_resize_bilinear.npz) and, where cv2 is importable,
    tests/test_oracle_crosscheck.py."""
    img = np.asarray(img, np.uint8)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    h, w = img.shape[:2]
    if (new_h, new_w) == (h, w):
        out = img.copy()
    elif h == 2 * new_h and w == 2 * new_w:
        a = img.astype(np.int32)
        out = ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    else:
        sx, fx = _linear_coeffs(new_w, w)
        lo, hi = sx < 0, sx >= w - 1
        fx = np.where(lo | hi, np.float32(0), fx).astype(np.float32)
        sx = np.clip(sx, 0, w - 1)
        a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int32)
        a1 = np.rint(fx * np.float32(2048)).astype(np.int32)
        sx1 = np.minimum(sx + 1, w - 1)                 # weight 0 where sx is the last column
        src = img.astype(np.int32)
        rows = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]        # [h, new_w, C]
        sy, fy = _linear_coeffs(new_h, h)
        b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int32)
        b1 = np.rint(fy * np.float32(2048)).astype(np.int32)
        r0 = rows[np.clip(sy, 0, h - 1)]
        r1 = rows[np.clip(sy + 1, 0, h - 1)]
        out = ((((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def pad_to_divisor(img, divisor=32, value=0):
    """transforms_disparity.py:177-194: bottom/right constant pad."""
    h, w = img.shape[:2]
    ph = int(np.ceil(h / divisor)) * divisor
    pw = int(np.ceil(w / divisor)) * divisor
    out = np.full((ph, pw) + img.shape[2:], value, dtype=img.dtype)
    out[:h, :w] = img
    return out


def disp2depth(disp, baseline=0.25, focal_length=640):
    """ocsort_disparity.py:132-134: Z = B*f / (d + 1e-6) in float32.

    torch evaluates ``python_scalar / tensor`` as ``tensor.reciprocal() *
    scalar`` (Tensor.__rtruediv__), i.e. two float32 roundings."""
    bf = np.float32(baseline * focal_length)
    return (np.float32(1) / (np.asarray(disp, np.float32) + np.float32(1e-6))) * bf


def preprocess(raw_u16, img_scale=None, size_divisor=32, baseline=0.25,
               focal_length=640):
    """Whole A1 chain for one frame.

    Returns dict(disp [H',W'] f32, mask [H',W'] u8, depth [H',W'] f32)."""
    disp, mask = load_postprocess(raw_u16)
    h, w = disp.shape
    if img_scale is not None:
        nh, nw = rescale_size(h, w, img_scale)
        disp = nearest_resize(disp, nh, nw)
        mask = nearest_resize(mask, nh, nw)
    disp = pad_to_divisor(disp, size_divisor, 0)
    mask = pad_to_divisor(mask, size_divisor, 0)
    return dict(disp=disp, mask=mask,
                depth=disp2depth(disp, baseline, focal_length))


def pack_image(img, size_divisor=32, pad_val=(114, 114, 114), mean=None, std=None, swap_rb=False, img_scale=None):
    """The RGB half of the detector input: Resize_Disparity's image branch when ``img_scale`` is
    given (keep_ratio; bilinear_resize_u8 to rescale_size), Pad_Disparity._pad_img
    (transforms_disparity.py:243 -> mmdet Pad: bottom/right pad of the HWC u8 image to a multiple
    of ``size_divisor`` with pad_val['img']), PackTrackInputs_Disparity HWC -> CHW
    (formatting_disparity.py:166-176), TrackDataPreprocessor_Disparity_V1.forward
    (data_preprocessor_disparity_v1.py:38-52): channel swap, float32, (img - mean) / std with
    mean / std in the swapped order, stack_batch (already at the padded size).
    img [H,W,3] u8 -> [3,H',W'] float32."""
    img = np.asarray(img, np.uint8)
    if img_scale is not None:
        img = bilinear_resize_u8(img, *rescale_size(img.shape[0], img.shape[1], img_scale))
    H, W = img.shape[:2]
    Hp = -(-H // size_divisor) * size_divisor
    Wp = -(-W // size_divisor) * size_divisor
    padded = np.empty((Hp, Wp, 3), np.uint8)
    padded[:] = np.asarray(pad_val, np.float64).astype(np.uint8)
    padded[:H, :W] = img
    chw = padded.transpose(2, 0, 1)
    if swap_rb:
        chw = chw[[2, 1, 0]]
    out = chw.astype(np.float32)
    if mean is not None:
        m = np.asarray(mean, np.float32).reshape(3, 1, 1)
        sd = np.asarray(std, np.float32).reshape(3, 1, 1)
        out = (out - m) / sd
    return np.ascontiguousarray(out, np.float32)
