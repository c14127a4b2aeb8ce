"""SE(3) helpers of the `is_6dof` deformation mode (reference: dgmesh/utils/rigid_utils.py:4-117; used by
`DeformNetwork*.forward` time_utils.py:116-123 and `render` gaussian_renderer/__init__.py:68-73).

Same functions, same argument meaning, same results to fp32 rounding -- written as closed-form elementwise
expressions (no 3x3 `bmm` chains): with W = [w]_x,
    exp_so3(w, theta) = I + sin(theta) W + (1 - cos(theta)) W^2
    exp_se3(S = (w, v), theta) = [[exp_so3(w, theta), G v], [0, 1]],
    G = theta I + (1 - cos(theta)) W + (theta - sin(theta)) W^2."""
import torch


def skew(w):
    """[N,3] -> [N,3,3] with skew(w) @ u == w x u."""
    x, y, z = w.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack([torch.stack([o, -z, y], -1), torch.stack([z, o, -x], -1), torch.stack([-y, x, o], -1)], -2)


def _skew_sq(w):
    """W^2 = w w^T - |w|^2 I."""
    return w.unsqueeze(-1) * w.unsqueeze(-2) - (w * w).sum(-1)[..., None, None] * torch.eye(3, device=w.device,
                                                                                          dtype=w.dtype)


def rp_to_se3(R, p):
    """R [N,3,3], p [N,3,1] -> homogeneous transforms [N,4,4]."""
    top = torch.cat([R, p], -1)
    bottom = torch.zeros_like(top[:, :1, :])
    bottom[..., 3] = 1.0
    return torch.cat([top, bottom], -2)


def exp_so3(w, theta):
    """Rodrigues: axis w [N,3], angle theta [N,1] -> rotation matrices [N,3,3]."""
    th = theta.reshape(-1, 1, 1)
    eye = torch.eye(3, device=w.device, dtype=w.dtype)
    return eye + torch.sin(th) * skew(w) + (1.0 - torch.cos(th)) * _skew_sq(w)


def exp_se3(S, theta):
    """Screw axis S = (w, v) [N,6], magnitude theta [N,1] -> [N,4,4]."""
    w, v = S[..., :3], S[..., 3:]
    th = theta.reshape(-1, 1, 1)
    eye = torch.eye(3, device=S.device, dtype=S.dtype)
    G = th * eye + (1.0 - torch.cos(th)) * skew(w) + (th - torch.sin(th)) * _skew_sq(w)
    return rp_to_se3(exp_so3(w, theta), G @ v.unsqueeze(-1))


def to_homogenous(v):
    return torch.cat([v, torch.ones_like(v[..., :1])], -1)


def from_homogenous(v):
    return v[..., :3] / v[..., -1:]
