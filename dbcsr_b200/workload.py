"""
dbcsr_b200/workload.py -- synthetic block-sparse matrices of the BASELINE.json configs (seeded, numpy PCG64).

Layout = what DBCSR hands to the local multiply: BCSR-ordered list index (row, col, blk_p) with 1-based coordinates and
element offsets, plus one flat data area of column-major blocks (src/core/dbcsr_types.F:340-461,499-526).
Block presence is i.i.d. Bernoulli(occupation) (SURVEY.md 8d allows this instead of the dlarnv geometric-skipping stream,
which lives in oracle/ and is used by the golden tests); values are uniform(0,1) like dlarnv(idist=1).
"""
import numpy as np


class Panel:
    def __init__(self, row_sizes, col_sizes, rows, cols, data=None, rng=None):
        self.row_sizes = np.ascontiguousarray(row_sizes, dtype=np.int32)
        self.col_sizes = np.ascontiguousarray(col_sizes, dtype=np.int32)
        self.rows = np.ascontiguousarray(rows, dtype=np.int32)  # 1-based, BCSR order
        self.cols = np.ascontiguousarray(cols, dtype=np.int32)
        nze = self.row_sizes[self.rows - 1].astype(np.int64) * self.col_sizes[self.cols - 1].astype(np.int64)
        self.offsets = np.zeros(self.rows.size, dtype=np.int64)
        if self.rows.size:
            self.offsets[1:] = np.cumsum(nze)[:-1]
        self.nze = int(nze.sum())
        if self.nze >= 2 ** 31:
            raise ValueError("data area exceeds int32 element offsets")
        if data is not None:
            self.data = data
        elif rng is not None:
            self.data = rng.random(self.nze)
        else:
            self.data = np.zeros(self.nze)

    @property
    def nblks(self):
        return int(self.rows.size)

    def list3(self):
        out = np.empty((self.nblks, 3), dtype=np.int32)
        out[:, 0], out[:, 1], out[:, 2] = self.rows, self.cols, self.offsets + 1
        return out

    def block(self, i):
        m, n = int(self.row_sizes[self.rows[i] - 1]), int(self.col_sizes[self.cols[i] - 1])
        return self.data[self.offsets[i]:self.offsets[i] + m * n].reshape(n, m).T  # col-major block as (m, n) view

    def sub(self, row_lo, row_hi, col_lo, col_hi):
        """Sub-panel with block rows in (row_lo, row_hi] and cols in (col_lo, col_hi], panel-local coordinates, own data area."""
        sel = np.nonzero((self.rows > row_lo) & (self.rows <= row_hi) & (self.cols > col_lo) & (self.cols <= col_hi))[0]
        p = Panel(self.row_sizes[row_lo:row_hi], self.col_sizes[col_lo:col_hi], self.rows[sel] - row_lo, self.cols[sel] - col_lo)
        nze = (self.row_sizes[self.rows[sel] - 1].astype(np.int64) * self.col_sizes[self.cols[sel] - 1].astype(np.int64))
        # gather block data (vectorised over equal block sizes is overkill here; panels are built once per run)
        idx = np.concatenate([np.arange(o, o + z) for o, z in zip(self.offsets[sel], nze)]) if sel.size else np.zeros(0, dtype=np.int64)
        p.data = self.data[idx]
        return p


def block_sizes(nblk, sizes, rng):
    sizes = list(sizes)
    if len(sizes) == 1:
        return np.full(nblk, sizes[0], dtype=np.int32)
    return rng.choice(np.array(sizes, dtype=np.int32), nblk).astype(np.int32)


def random_panel(row_sizes, col_sizes, occupation, rng):
    nr, nc = len(row_sizes), len(col_sizes)
    mask = rng.random((nr, nc)) < occupation
    rows, cols = np.nonzero(mask)  # row-major => BCSR order
    return Panel(row_sizes, col_sizes, rows + 1, cols + 1, rng=rng)


def make_config(name, seed=42, nblk=None):
    """BASELINE.json configs: 'cfg2' = 23x23 FP64, N=1000 block rows/cols/k, 10 % occupation (configs[0..1], [4]);
    'cfg3' = mixed {5,13,23,26,32}, 5 %; 'cfg4' = 23x23, 50 % (BF16 config). nblk overrides N (parity tests use small N)."""
    rng = np.random.default_rng(seed)
    if name == "cfg2":
        n = nblk or 1000
        sizes, occ = [23], 0.10
    elif name == "cfg3":
        n = nblk or 1000
        sizes, occ = [5, 13, 23, 26, 32], 0.05
    elif name == "cfg4":  # 23x23, 50 % occupation, BF16 operands / FP32 C on the tensor cores (extension dtype)
        n = nblk or 1000
        sizes, occ = [23], 0.50
    else:
        raise ValueError(name)
    bs = block_sizes(n, sizes, rng)  # same size vector for rows, cols and k
    A = random_panel(bs, bs, occ, rng)
    B = random_panel(bs, bs, occ, rng)
    return dict(name=name, nblk=n, sizes=sizes, occupation=occ, m_sizes=bs, n_sizes=bs, k_sizes=bs, A=A, B=B, seed=seed)
