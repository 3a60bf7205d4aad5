#!/usr/bin/env python3
"""Shrink a real pyngp-trained snapshot into a committable fixture WITHOUT changing what it renders.

The renderer only consumes the density grid through the occupancy bitfield
(`grid > min(0.01, mean(max(grid,0)))`, reference src/testbed_nerf.cu:284-312, 2355-2373).
Replacing every cell by +1 (occupied) / -1 (empty) keeps every bit identical as long as more
than 1 % of cascade 0 is occupied (then the threshold stays 0.01) -- asserted below -- and
compresses 8 MB of noisy fp16 to ~100 KB.  Weights are untouched.

    python tests/golden/repack_fixture.py gpurun_out/golden/fox_a2_small.ingp tests/golden/fox_a2_small_packed.ingp
"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
from dream2real_b200 import ingp  # noqa: E402
from oracle import ngp_oracle as O  # noqa: E402

src, dst = sys.argv[1], sys.argv[2]
snap = ingp.load_snapshot(src)
bits0, thr0 = O.build_bitfield(snap.density_grid, snap.max_cascade)
occ = snap.density_grid > np.float32(thr0)
packed = np.where(occ, np.float16(1.0), np.float16(-1.0)).astype(np.float16)
bits1, thr1 = O.build_bitfield(packed.astype(np.float32), snap.max_cascade)
assert np.array_equal(bits0, bits1), "bitfield changed"
cfg = snap.config
cfg["snapshot"]["density_grid_binary"] = packed.tobytes()
# per-image optimiser state is irrelevant for rendering; keep one entry per image but drop nothing else
ingp.save_snapshot(dst, cfg, compress_level=9)
again = ingp.load_snapshot(dst)
assert np.array_equal(again.params, snap.params)
assert np.array_equal(O.build_bitfield(again.density_grid, again.max_cascade)[0], bits0)
print("ok", dst, "threshold", thr0, "->", thr1)
