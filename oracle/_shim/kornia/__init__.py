"""Import stand-in for `kornia` (not installed in this image), used ONLY by oracle/make_golden.py so that the
reference's `gluefactory.geometry.gt_generation` can be imported: that module reaches kornia through
`geometry/depth.py:1`, but the function pinned here (`gt_matches_from_homography`, gt_generation.py:109-161) never
calls it.  Any attribute access fails loudly."""


def __getattr__(name):
    raise ImportError(f"kornia stand-in: '{name}' is not available (kornia is not installed)")
