"""The C-ABI library loads and exports every symbol include/*.h declares (no
compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    from stereotracking_b200 import build
    return build.build()


def test_header_symbols_are_exported(built):
    header = open(os.path.join(ROOT, 'include', 'stereotrack_b200.h')).read()
    declared = set(re.findall(r'\b(st_[a-z0-9_]+)\s*\(', header))
    assert {'st_create', 'st_disp2depth', 'st_decode_nms', 'st_box_depth', 'st_track_step'} <= declared
    lib = ctypes.CDLL(built)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    from stereotracking_b200 import _lib
    assert declared == set(_lib.exported_symbols())


def test_abi_version_and_errors(built):
    from stereotracking_b200 import _lib
    lib = _lib.lib()
    assert lib.st_abi_version() == _lib.ABI_VERSION
    assert lib.st_error_string(0) == b'ok'
    assert b'invalid' in lib.st_error_string(-1)
    cfg = _lib.StConfig()
    lib.st_default_config(ctypes.byref(cfg))
    # constructor defaults of the reference classes
    assert abs(cfg.obj_score_thr - 0.3) < 1e-7 and abs(cfg.match_iou_thr - 0.3) < 1e-12
    assert cfg.num_tentatives == 3 and cfg.vel_delta_t == 3 and cfg.num_frames_retain == 10
    assert cfg.weight_iou_with_det_scores == 1


def test_config_struct_matches_header_layout(built):
    """st_config field order in _lib.py must match the header."""
    header = open(os.path.join(ROOT, 'include', 'stereotrack_b200.h')).read()
    body = header[header.index('typedef struct st_config {'):header.index('} st_config;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = re.findall(r'(?:int32_t|float|double)\s+([a-z_]+);', body)
    from stereotracking_b200 import _lib
    assert fields == [f[0] for f in _lib.StConfig._fields_]


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from stereotracking_b200 import _lib
    from stereotracking_b200.engine import StereoTrackEngine
    with pytest.raises(_lib.StereoTrackError):
        StereoTrackEngine()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'stereotracking_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert 'oracle/' not in src or f == 'build.py', f
