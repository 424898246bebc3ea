"""Small O(M+N) geometry helpers of the ground-truth path (torch elementwise ops, any device).

`warp_points` restates geometry/homography.py:161-180 (`warp_points_torch`: homogeneous multiply by H^T, divide with the
eps of geometry/utils.py:22-30); `inv3x3` replaces the `torch.inverse` of :176 with the closed-form adjugate evaluated
in fp64 (elementwise only, so it can be captured in a CUDA graph -- torch.inverse synchronises inside cuSOLVER -- and
more accurate than an fp32 LU; the labels are identical on the reference-generated goldens).
"""
import torch


def warp_points(pts, Hm):
    """pts [B,N,2] -> H . pts (homogeneous divide, eps as geometry/utils.from_homogeneous)."""
    ones = torch.ones_like(pts[..., :1])
    ph = torch.cat([pts, ones], -1) @ Hm.transpose(-1, -2)
    return ph[..., :2] / (ph[..., 2:] + 1e-5)


def inv3x3(H):
    """Inverse of [..., 3, 3] matrices: adjugate / determinant in fp64, rounded to fp32."""
    Hd = H.double()
    a, b, c, d, e, f, g, h, i = (Hd[..., r, k] for r in range(3) for k in range(3))
    A, Bc, C = e * i - f * h, -(d * i - f * g), d * h - e * g
    det = a * A + b * Bc + c * C
    adj = torch.stack([A, -(b * i - c * h), b * f - c * e, Bc, a * i - c * g, -(a * f - c * d), C, -(a * h - b * g),
                       a * e - b * d], -1).reshape(Hd.shape)
    return (adj / det[..., None, None]).float()
