"""Row-vector 4x4 transforms (restatement of pytorch3d.transforms.Transform3d semantics).

Points are row vectors: ``p_out = [p, 1] @ M`` followed by a divide by the 4th component.
``compose`` keeps a list of matrices that are multiplied left-to-right when the matrix is
requested; ``inverse`` inverts every factor separately and reverses the list (this is how
PyTorch3D avoids one big ``torch.inverse`` of the composed matrix, and it fixes the fp32
rounding sequence of ray un-projection).
"""
import torch


def _bmm(a, b):
    """Batched matmul with broadcasting of a leading batch dim of size 1."""
    if a.dim() == 2:
        a = a[None]
    if b.dim() == 2:
        b = b[None]
    if a.shape[0] != b.shape[0]:
        if a.shape[0] == 1:
            a = a.expand(b.shape[0], -1, -1)
        elif b.shape[0] == 1:
            b = b.expand(a.shape[0], -1, -1)
        else:
            raise ValueError("Expected batch dim for bmm to be equal or 1; got %r, %r" % (a.shape, b.shape))
    return a.bmm(b)


class Transform3d:
    def __init__(self, matrix=None, dtype=torch.float32, device="cpu", factors=None, inv_fn=None):
        if matrix is None:
            matrix = torch.eye(4, dtype=dtype, device=device)[None]
        if matrix.dim() == 2:
            matrix = matrix[None]
        self._matrix = matrix
        self._factors = list(factors) if factors else []  # transforms applied AFTER self._matrix
        self._inv_fn = inv_fn                              # cheap inverse of self._matrix, if known
        self.device = matrix.device
        self.dtype = matrix.dtype

    # -- construction helpers -------------------------------------------------------------
    @staticmethod
    def rotate(R):
        """R: [N,3,3] applied as p @ R; inverse is the transpose."""
        M = torch.zeros(R.shape[0], 4, 4, dtype=R.dtype, device=R.device)
        M[:, :3, :3] = R
        M[:, 3, 3] = 1.0
        return Transform3d(M, inv_fn=lambda m: m.permute(0, 2, 1).contiguous())

    @staticmethod
    def translate(t):
        """t: [N,3]; inverse negates the translation row."""
        M = torch.eye(4, dtype=t.dtype, device=t.device)[None].repeat(t.shape[0], 1, 1)
        M[:, 3, :3] = t

        def inv(m):
            out = m.clone()
            out[:, 3, :3] = -m[:, 3, :3]
            return out
        return Transform3d(M, inv_fn=inv)

    @staticmethod
    def scale(s):
        """s: [N,3]; inverse is 1/s on the diagonal."""
        M = torch.zeros(s.shape[0], 4, 4, dtype=s.dtype, device=s.device)
        M[:, 0, 0], M[:, 1, 1], M[:, 2, 2] = s[:, 0], s[:, 1], s[:, 2]
        M[:, 3, 3] = 1.0

        def inv(m):
            d = torch.diagonal(m, dim1=1, dim2=2)
            return torch.diag_embed(1.0 / d)
        return Transform3d(M, inv_fn=inv)

    # -- algebra ---------------------------------------------------------------------------
    def compose(self, *others):
        return Transform3d(self._matrix, factors=self._factors + list(others), inv_fn=self._inv_fn)

    def get_matrix(self):
        M = self._matrix
        for f in self._factors:
            M = _bmm(M, f.get_matrix())
        return M

    def _own_inverse(self):
        if self._inv_fn is not None:
            return self._inv_fn(self._matrix)
        return torch.inverse(self._matrix)

    def inverse(self):
        own = Transform3d(self._own_inverse())
        if not self._factors:
            return own
        ident = Transform3d(dtype=self.dtype, device=self.device)
        return ident.compose(*[f.inverse() for f in reversed(self._factors)], own)

    def transform_points(self, points, eps=None):
        pts = points
        if pts.dim() == 2:
            pts = pts[None]
        ones = torch.ones(pts.shape[0], pts.shape[1], 1, dtype=pts.dtype, device=pts.device)
        homog = torch.cat([pts, ones], dim=2)
        out = _bmm(homog, self.get_matrix())
        denom = out[..., 3:]
        if eps is not None:
            sign = denom.sign() + (denom == 0.0).type_as(denom)
            denom = sign * torch.clamp(denom.abs(), eps)
        out = out[..., :3] / denom
        if out.shape[0] == 1 and points.dim() == 2:
            out = out.reshape(points.shape)
        return out

    def to(self, device):
        return Transform3d(self._matrix.to(device), factors=[f.to(device) for f in self._factors],
                           inv_fn=self._inv_fn)
