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
SOURCES = ['st_context.cu', 'k1_disp2depth.cu', 'k1b_pack_image.cu', 'k2_decode_nms.cu', 'k3_box_depth.cu', 'k3_box_depth_raw.cu', 'k4_track.cu', 'k4_track2.cu', 'k_pipeline.cu']
HEADERS = ['st_common.cuh', 'kalman.cuh', 'lapjv.cuh', 'k3_common.cuh', 'k3_select.cuh', 'k4_common.cuh', os.path.join('..', '..', 'include', 'stereotrack_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-fmad=false', '--extended-lambda', '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


OBJ_DIR = os.path.join(HERE, 'build')
COMPILE_FLAGS = [f for f in NVCC_FLAGS if f != '-shared']


def _deps():
    return [os.path.join(CSRC, f) for f in HEADERS]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(args):
    src, obj, verbose = args
    cmd = [_nvcc()] + COMPILE_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr, cmd


def build(force=False, verbose=False):
    """One object per translation unit (compiled in parallel, only the stale ones), then the link."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_time = max(os.path.getmtime(d) for d in _deps())
    jobs, objs = [], []
    for f in SOURCES:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ_DIR, f[:-3] + '.o')
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append((src, obj, verbose))
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for src, rc, out, cmd in pool.map(_compile_one, jobs):
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
            if verbose:
                sys.stderr.write(out)
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('link failed: ' + ' '.join(cmd))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
