"""Compile the CUDA kernels + C ABI into ``libstereotrack_b200.so`` (in-tree).

``python -m stereotracking_b200.build`` or ``__graft_entry__.build()``.
nvcc cross-compiles for sm_100a without a GPU.  ``-fmad=false`` is part of the
numerical contract (DESIGN.md "Numerics"): float32/float64 expressions must
round like the reference's un-fused NumPy/torch arithmetic.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libstereotrack_b200.so')
SOURCES = ['k1_disp2depth.cu', 'k1b_pack_image.cu', 'k2_decode_nms.cu', 'k3_box_depth.cu', 'k4_track.cu', 'k_pipeline.cu']
HEADERS = ['st_common.cuh', 'kalman.cuh', 'lapjv.cuh', os.path.join('..', '..', 'include', 'stereotrack_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-fmad=false', '--extended-lambda', '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
        ['-o', LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
