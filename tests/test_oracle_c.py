"""Two independent restatements of the decode (numpy op-by-op and plain C loops) must agree."""
import ctypes as C
import os
import subprocess

import numpy as np

from metro_pose3d_b200.joints import export_permutation
from metro_pose3d_b200.weights import synth_head
from oracle.metro_oracle import decode_ref

ORACLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle')


def test_c_and_numpy_oracles_agree():
    subprocess.run(['make', '-C', ORACLE, '-s'], check=True)
    lib = C.CDLL(os.path.join(ORACLE, '_build', 'libdecode_ref.so'))
    for side, stride, j, ds in [(8, 32, 17, 'h36m'), (16, 16, 19, 'coco19'), (8, 32, 53, 'merged')]:
        perm = export_permutation(ds)
        x = synth_head(3, side, j, seed=side)
        out = np.zeros((3, len(perm), 3))
        p = (C.c_int * len(perm))(*perm)
        rc = lib.metro_oracle_decode(x.ctypes.data_as(C.c_void_p), 3, side, j, 8, stride, 1, 256, C.c_double(2200.0),
                                     p, len(perm), out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert np.abs(out - decode_ref(x, j, stride, perm)).max() < 1e-9
