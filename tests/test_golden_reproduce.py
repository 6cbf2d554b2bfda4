"""The committed fixtures ARE what the unmodified reference produces: where the reference tree
is present (the build container), regenerate every fixture with tests/golden/make_golden.py into
a scratch directory and compare array by array.  Skipped on the GPU box (no /root/reference)."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN

REF = os.environ.get('STEREOTRACKING_REF', '/root/reference')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'mmtrack')), reason='reference tree not present')
def test_committed_fixtures_are_regenerated_bit_for_bit(tmp_path):
    env = dict(os.environ, GOLDEN_OUT=str(tmp_path))
    res = subprocess.run([sys.executable, os.path.join(GOLDEN, 'make_golden.py')], env=env, capture_output=True,
                         text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    committed = sorted(glob.glob(os.path.join(GOLDEN, '*.npz')))
    assert len(committed) >= 12
    for path in committed:
        fresh = np.load(os.path.join(str(tmp_path), os.path.basename(path)))
        kept = np.load(path)
        assert set(fresh.files) == set(kept.files), path
        for k in kept.files:
            assert fresh[k].dtype == kept[k].dtype and fresh[k].shape == kept[k].shape, (path, k)
            assert np.array_equal(fresh[k], kept[k], equal_nan=kept[k].dtype.kind == 'f'), (path, k)
