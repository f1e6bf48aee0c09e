"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the region scan in the reference's query generator,
catkin_ws/src/local_navigator/src/local_navigator.py:81-257 (callbackOccGrid).  It is the checker of
the CUDA kernel behind pc_front_regions; the product never imports it.  Pinned against the reference
module itself (oracle/ref_navigator.py) by tests/test_navigator_ref_pin.py.

Bit layout = PC_REGION_* of include/prius_costmap.h.
"""
import numpy as np

BANDS = ((180, 209), (210, 239), (240, 269), (270, 299))      # y ranges of F1/R1/L1 .. F4/R4/L4 (:120-257)
F_X, R_X, L_X = (140, 159), (120, 139), (160, 179)


def region_mask(data, info_width):
    """data: flat int8 sequence of the OccupancyGrid, info_width: msg.info.width (:82-83 use it twice)."""
    w = int(info_width)
    board = np.asarray(data)[: w * w].reshape(w, w)            # board[x][y] = data[x * w + y]  (:101-110)
    mask, any_f, any_l = 0, False, False
    for k, (y0, y1) in enumerate(BANDS):
        f = board[F_X[0]:F_X[1] + 1, y0:y1 + 1]                # slices clip to the grid like the loops do
        if (f == 100).any():                                    # F_k: any hit (:120-150)
            mask |= 1 << k
            any_f = True
        for base, seen, (x0, x1) in ((4, 16, R_X), (8, 20, L_X)):
            r = board[x0:x1 + 1, y0:y1 + 1]
            if r.size:                                          # R_k / L_k: the LAST cell scanned decides (:155-257)
                mask |= 1 << (seen + k)
                if r[-1, -1] == 100:
                    mask |= 1 << (base + k)
                if base == 8 and (r == 100).any():
                    any_l = True
    # self.obstacle: x ascends, so R cells (write False) come first, F cells (write True) next and
    # L cells (write False) last
    if any_l:
        mask |= 1 << 13
    if any_f and not any_l:
        mask |= 1 << 12
    return mask
