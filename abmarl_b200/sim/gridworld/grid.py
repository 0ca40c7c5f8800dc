"""Grid declaration -- mirrors abmarl/sim/gridworld/grid.py:7-71.

The reference Grid is also the occupancy *store* (an object array of insertion-ordered dicts); in this
engine the store is device state (BgwState.cell/next, include/bgw.h), so this class keeps only the shape and
the symmetrised overlapping map that every env of the batch shares.
"""


class Grid:
    def __init__(self, rows, cols, overlapping=None, **kwargs):
        assert type(rows) is int and rows > 0, "Rows must be a positive integer."
        assert type(cols) is int and cols > 0, "Cols must be a positive integer."
        self.rows, self.cols = rows, cols
        self.overlapping = overlapping

    @property
    def overlapping(self):
        return self._overlapping

    @overlapping.setter
    def overlapping(self, value):
        if value is None:
            self._overlapping = {}
            return
        assert type(value) is dict, "Overlaping must be dictionary."
        closed = {k: set(v) for k, v in value.items()}
        for enc, partners in value.items():
            assert type(enc) is int, "All keys in overlapping dict must be integers."
            assert type(partners) is set, "All values in overlapping dict must be sets."
            for other in partners:
                assert type(other) is int, "All elements in overlapping dict values must be integers."
                closed.setdefault(other, set()).add(enc)      # 2 overlaps 3  =>  3 overlaps 2 (grid.py:64-68)
        self._overlapping = closed
