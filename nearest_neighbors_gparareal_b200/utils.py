"""Affine state normalisation used by the ODE systems (interface of reference utils.py:1-33) and the
dimension partition of the multi-GPU sweep.

`Normalize('-11')` maps the box [mn, mx] onto [-1, 1]^d; on the device the same map is fused into the
right-hand sides (csrc/rk.cu), this host object serves initial conditions and tests.
"""
_KINDS = ('identity', '-11')


class Normalize():
    """x -> 2 (x - mn) / (mx - mn) - 1 for kind '-11', the identity otherwise.

    Methods and error text follow the reference class so that ODE subclasses written against it work
    unchanged: fit (forward map), inverse, get_scale (derivative of the forward map)."""

    def __init__(self, mn, mx, norm_type=None):
        kind = 'identity' if norm_type is None else norm_type.lower()
        if kind not in _KINDS:
            raise NotImplementedError('Only identity and -11 are implemented')
        self.norm_type = kind
        self.mn, self.mx = mn, mx

    @property
    def _span(self):
        return self.mx - self.mn

    def fit(self, x):
        return x if self.norm_type == 'identity' else 2 * (x - self.mn) / self._span - 1

    def inverse(self, x):
        return x if self.norm_type == 'identity' else (x + 1) / 2 * self._span + self.mn

    def get_scale(self):
        return 1 if self.norm_type == 'identity' else 2 / self._span


def dim_block(d, rank, world):
    """Output dimensions [j0, j0+dl) whose fits `rank` runs when a predict is sharded by dimension over
    `world` ranks (equal blocks: d must be a multiple of world)."""
    dl = d // world
    return rank * dl, dl
