"""CPU oracle for the StereoTracking depth-guided tracking path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product path
(``stereotracking_b200``) never imports, links or executes anything in this
directory and fails loudly when its CUDA library is missing.

What it is: a NumPy/SciPy (+ a small C file for the LAP solver) restatement of
the reference algorithm, each function citing the reference file:line it
follows (paths relative to the StereoTracking repository root).

Pinning status (see DESIGN.md section "Oracle"):
  * tracker / Kalman filter / per-box depth / box scaling:
    PINNED against the unmodified reference modules executed in the build
    container through ``tests/golden/make_golden.py`` (third-party imports
    replaced by named stand-ins); the resulting vectors are committed under
    ``tests/golden/`` and replayed by ``tests/test_oracle_golden.py``.
  * ``lap.lapjv`` (third-party, unpinned in requirements/runtime.txt:2),
    mmdet ``bbox_overlaps``, the mmyolo/mmdet/mmcv head post-process + NMS and
    mmcv ``imrescale`` are NOT in the reference tree: restated from their
    published algorithms and cross-checked against SciPy
    ``linear_sum_assignment``, ``torchvision.ops.box_iou`` / ``nms`` and
    ``cv2.resize`` -> for those rows parity is "unpinned by the reference".

Floating-point conventions shared with the CUDA path (so that index outputs
can be compared bit-exactly): float32 +,-,*,/ and sqrt are IEEE and evaluated
in the reference's order without FMA contraction; exp / sigmoid / acos are
evaluated in float64 and rounded once to float32 (the correctly rounded
value, <= 1 ulp from either torch backend the reference may run on).
"""
from . import disparity, head, boxdepth, kalman, lapjv, tracker, model  # noqa: F401
