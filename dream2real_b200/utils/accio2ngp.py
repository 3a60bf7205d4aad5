"""Frame conversion used on the path (reference utils/accio2ngp.py:133-139)."""
import numpy as np


def converter(T_accio_list):
    """OpenCV-style camera/object poses -> NeRF convention: negate the y and z rotation columns.
    Accepts [M,4,4] (or a single [4,4]); returns a copy like the reference does."""
    T = np.array(T_accio_list, copy=True)
    T[..., :3, 1] *= -1
    T[..., :3, 2] *= -1
    return T
