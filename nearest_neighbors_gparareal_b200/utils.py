"""Host mirror of reference utils.py:1-33 (`Normalize`)."""


class Normalize():
    def __init__(self, mn, mx, norm_type=None):
        self.mn = mn
        self.mx = mx
        if norm_type is None:
            norm_type = 'identity'
        if norm_type.lower() not in ['identity', '-11']:
            raise NotImplementedError('Only identity and -11 are implemented')
        self.norm_type = norm_type.lower()

    def fit(self, x):
        if self.norm_type == '-11':
            return 2 * (x - self.mn) / (self.mx - self.mn) - 1
        return x

    def inverse(self, x):
        if self.norm_type == '-11':
            return (x + 1) / 2 * (self.mx - self.mn) + self.mn
        return x

    def get_scale(self):
        if self.norm_type == '-11':
            return 2 / (self.mx - self.mn)
        return 1


def dim_block(d, rank, world):
    """Output dimensions [j0, j0+dl) whose fits `rank` runs when a predict is sharded by dimension over
    `world` ranks (equal blocks: d must be a multiple of world)."""
    dl = d // world
    return rank * dl, dl
