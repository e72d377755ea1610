"""Drop-in for the reference's ThinPlateSpline2.py (/root/reference/ThinPlateSpline2.py:4,167-170):
identical to ThinPlateSpline except that the right-hand side of the TPS system is the absolute
`target` positions instead of `coord + vector` (ThinPlateSpline2.py:160)."""
from . import ops


def ThinPlateSpline2(U, source, target, out_size, return_grid=True):
    """U [B,H,W,C]; source, target [B, num_point, 2]; out_size (h, w) -> (output, x, y)."""
    return ops.thin_plate_spline(U, source, target, out_size, want_grid=return_grid)
