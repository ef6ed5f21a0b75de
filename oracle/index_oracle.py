"""
oracle/index_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pure-Python restatement of the integer (index) work on DBCSR's local-multiply path, used to check the
product's C++ stack builder (dbcsr_b200/csrc/host) entry by entry on small cases:

  rec_sort_index / rec_split   src/mm/dbcsr_mm_common.F:227-309
  sparse_multrec, find_cut_*   src/mm/dbcsr_mm_multrec.F:487-658
  build_csr_index              src/mm/dbcsr_mm_csr.F:741-795
  dbcsr_mm_csr_multiply_low    src/mm/dbcsr_mm_csr.F:178-359
  dbcsr_mm_csr_init stack map  src/mm/dbcsr_mm_csr.F:361-538
  fill_hash_tables (preset C)  src/mm/dbcsr_mm_csr.F:540-576
  multrec_filtering            src/mm/dbcsr_mm_multrec.F:700-758
  row_max_epss, norm filter    src/mm/dbcsr_mm_cannon.F:1098-1107, src/mm/dbcsr_mm_csr.F:270-278
  flush_stacks / purge         src/mm/dbcsr_mm_csr.F:696-739
  map_most_common              src/dist/dbcsr_dist_util.F:753-812
  stack_sort / stack_binning   src/mm/dbcsr_mm_accdrv.F:364-423
  sort (stable merge sort)     src/utils/dbcsr_array_sort.F

Parity status: UNPINNED against a running reference (no Fortran compiler in the build container, no golden
index dumps in the reference's tests); the restatement follows the cited lines statement by statement.
All indices are 1-based like the Fortran; deliberately slow and simple (plain loops).
"""
import numpy as np


# --------------------------------------------------------------------------- rec_sort_index

def checker_tr(row, column):
    """src/dist/dbcsr_dist_operations.F:65-75 (copied to src/mm/dbcsr_mm_csr.F:797-810): BTEST(column + row, 0) .EQV. column >= row"""
    return bool((column + row) & 1) == (column >= row)

def rec_split(a, row_or_col, mi, half):
    """src/mm/dbcsr_mm_common.F:283-309: low part in order, high part filled from the END (i.e. reversed)."""
    nele = len(a)
    half_m = mi + half - 1
    split = [None] * nele
    p_low, p_high = 0, nele - 1
    for el in a:
        if el[row_or_col] <= half_m:
            split[p_low] = el
            p_low += 1
        else:
            split[p_high] = el
            p_high -= 1
    return split, p_low


def rec_sort_index(mi, mf, ni, nf, a):
    """src/mm/dbcsr_mm_common.F:227-281.  a: list of (row, col, blk_p) tuples. Returns the sorted list."""
    nele = len(a)
    M = mf - mi + 1
    N = nf - ni + 1
    if M > N:
        half = M // 2
        a, nlow = rec_split(a, 0, mi, half)
        lo, hi = a[:nlow], a[nlow:]
        if nlow > 1:
            lo = rec_sort_index(mi, mi + half - 1, ni, nf, lo)
        if nele - nlow > 1:
            hi = rec_sort_index(mi + half, mf, ni, nf, hi)
    else:
        half = N // 2
        a, nlow = rec_split(a, 1, ni, half)
        lo, hi = a[:nlow], a[nlow:]
        if nlow > 1:
            lo = rec_sort_index(mi, mf, ni, ni + half - 1, lo)
        if nele - nlow > 1:
            hi = rec_sort_index(mi, mf, ni + half, nf, hi)
    return lo + hi


# --------------------------------------------------------------------------- helpers
def stable_sort_index(keys):
    """dbcsr sort(): stable ascending (merge + bubble, strict '<' comparisons) -> 0-based permutation."""
    return sorted(range(len(keys)), key=lambda i: keys[i])


def map_most_common(array, nmost_common):
    """src/dist/dbcsr_dist_util.F:753-812. Returns (map list indexed by size 0..max_val, most_common_elements, max_val)."""
    if len(array) > 0:
        max_val = max(array)
        max_val_l = max_val
    else:
        max_val = 0
        max_val_l = 0
    size_counts = [0] * (max_val_l + 1)
    for v in array:
        if v <= max_val_l:
            size_counts[v] -= 1
    perm = stable_sort_index(size_counts) if len(array) > 0 else [0] * (max_val_l + 1)
    nmc = min(nmost_common, max_val_l)
    mc_map = [nmost_common + 1] * (max_val_l + 1)
    # Fortran: permutation holds 1-based positions of a 0-based-indexed array -> size = perm - 1; here perm is 0-based
    for i in range(1, nmc + 1):
        mc_map[perm[i - 1]] = i
    elements = [0] * nmost_common
    for i in range(nmc):
        elements[i] = perm[i]
    return mc_map, elements, max_val


def find_cut(lst, ai, af, which, val):
    """find_cut_row (which=0) / find_cut_col (which=1), src/mm/dbcsr_mm_multrec.F:579-658. 1-based ai..af."""
    ilow = ai
    if lst[ilow - 1][which] > val:
        return ilow
    ihigh = af
    if lst[ihigh - 1][which] <= val:
        return ihigh + 1
    while True:
        if ihigh - ilow == 1:
            break
        i = (ilow + ihigh) // 2
        if lst[i - 1][which] > val:
            ihigh = i
        else:
            ilow = i
    return ihigh


def row_max_epss(filter_eps, total_row_counts):
    """src/mm/dbcsr_mm_cannon.F:1098-1107: (filter_eps_sp / REAL(MAX(1, count)))**2, all in single precision."""
    eps_sp = np.float32(filter_eps)
    out = np.empty(len(total_row_counts), dtype=np.float32)
    for r, c in enumerate(total_row_counts):
        q = np.float32(eps_sp / np.float32(max(1, int(c))))
        out[r] = np.float32(q * q)
    return out


def block_norms(list3, row_sizes, col_sizes, data):
    """calc_norms_d, src/mm/dbcsr_mm_common.F:700-730: norms(blk) = REAL(SUM(DATA(bp:bpe)**2), KIND=sp) -- the SQUARED
    Frobenius norm accumulated in double, rounded to single.  list3: (row, col, blk_p) 1-based."""
    out = np.zeros(len(list3), dtype=np.float32)
    for i, (row, col, bp) in enumerate(list3):
        if bp != 0:
            nze = row_sizes[row - 1] * col_sizes[col - 1]
            blk = np.asarray(data[abs(bp) - 1:abs(bp) - 1 + nze], dtype=np.float64)
            out[i] = np.float32(np.sum(blk * blk))
    return out


def multrec_filtering(filter_eps, rowi, coli, blkp, rbs, cbs, data):
    """multrec_filtering_d, src/mm/dbcsr_mm_multrec.F:700-758: keep block iff DDOT(blk, blk) >= filter_eps**2 (double); kept
    entries move to the front in order, blk_p unchanged.  Returns (rowi, coli, blkp, nze) of the kept blocks and the norms used."""
    eps_opt = float(filter_eps) ** 2
    out_r, out_c, out_p, nze, norms = [], [], [], 0, []
    for r, c, bp in zip(rowi, coli, blkp):
        blk_nze = int(rbs[r - 1]) * int(cbs[c - 1])
        if bp == 0 or blk_nze == 0:
            norms.append(0.0)
            continue
        blk = np.asarray(data[bp - 1:bp - 1 + blk_nze], dtype=np.float64)
        nrm = float(np.dot(blk, blk))
        norms.append(nrm)
        if nrm >= eps_opt:
            out_r.append(int(r))
            out_c.append(int(c))
            out_p.append(int(bp))
            nze += blk_nze
    return out_r, out_c, out_p, nze, norms


def build_csr_index(mi, mf, ai, af, lst):
    """src/mm/dbcsr_mm_csr.F:741-795. Returns row_p dict-like list (offset by mi) and blk_info list of (col, blk_p)
    (plus the list position of each CSR entry as a third component, which is how csr_norms(:) follows list_norms(:))."""
    counts = [0] * (mf - mi + 1)
    for i in range(ai, af + 1):
        counts[lst[i - 1][0] - mi] += 1
    row_p = [0] * (mf - mi + 2)
    for r in range(1, mf - mi + 2):
        row_p[r] = row_p[r - 1] + counts[r - 1]
    blk_info = [None] * (af - ai + 1)
    counts = [0] * (mf - mi + 1)
    for i in range(ai, af + 1):
        row = lst[i - 1][0]
        counts[row - mi] += 1
        blk_info[row_p[row - mi] + counts[row - mi] - 1] = (lst[i - 1][1], lst[i - 1][2], i)
    return row_p, blk_info


class LocalMultiplyOracle:
    """One thread's multrec + csr + stack flush + accdrv sort, for one Cannon tick on one rank.

    a_list/b_list: BCSR-ordered lists of (row, col, blk_p) with 1-based LOCAL row/col and 1-based element
    offsets (blk_p > 0, i.e. blocks stored untransposed).  m_sizes/n_sizes/k_sizes: block sizes per local
    row of C / col of C / k.  Collects every dispatched stack in self.dispatched as a dict with
    'm','n','k' (descriptor values, 0 when inhomogeneous), 'defined_mnk', 'host' (S x 7 int32, original order)
    and 'dev' (S x 3 int32 after stack_sort / stack_binning).
    """

    def __init__(self, m_sizes, n_sizes, k_sizes, mm_stack_size=30000, n_stacks=3, multrec_limit=512,
                 stack_sort=True, min_flop_sort=4000, binning_nbins=4096, binning_binsize=16):
        self.m_sizes, self.n_sizes, self.k_sizes = list(m_sizes), list(n_sizes), list(k_sizes)
        self.mm_stack_size = mm_stack_size
        self.multrec_limit = multrec_limit
        self.cfg_sort, self.min_flop_sort = stack_sort, min_flop_sort
        self.nbins, self.binsize = binning_nbins, binning_binsize
        self.nn = self.nk = self.nm = n_stacks
        self.nstacks = n_stacks ** 3 + 1
        self._init_stack_map()
        self.stacks = [[] for _ in range(self.nstacks)]  # host 7-tuples per stack (index 0 = stack 1)
        # product work matrix
        self.c_row_i, self.c_col_i, self.c_blk_p = [], [], []
        self.datasize = 0
        self.c_hash = {}
        self.flop = 0
        self.dispatched = []
        # on-the-fly filter (use_eps): thresholds per C row and norms aligned with the SORTED lists (set by multiply)
        self.row_eps = None
        self.a_norms = self.b_norms = None
        self.skipped = 0
        self.keep_sparsity = False
        # product with symmetry (src/mm/dbcsr_mm_csr.F:280-292): global block index of the local C rows / cols, or None
        self.c_has_symmetry = False
        self.c_local_rows = self.c_local_cols = None

    def set_c_symmetry(self, on, global_rows=None, global_cols=None):
        self.c_has_symmetry = bool(on)
        self.c_local_rows, self.c_local_cols = global_rows, global_cols

    def preset_c(self, rows, cols, keep_sparsity=False):
        """Work matrix starts from existing C blocks in list order (fill_hash_tables, src/mm/dbcsr_mm_csr.F:540-576; offsets are
        the running sum of the block sizes); keep_sparsity = retain_sparsity of dbcsr_multiply (src/mm/dbcsr_mm_csr.F:307)."""
        for r, c in zip(rows, cols):
            self.c_row_i.append(int(r))
            self.c_col_i.append(int(c))
            self.c_blk_p.append(self.datasize + 1)
            self.datasize += int(self.m_sizes[r - 1]) * int(self.n_sizes[c - 1])
            self.c_hash[(int(r), int(c))] = len(self.c_blk_p)
        self.keep_sparsity = keep_sparsity

    # src/mm/dbcsr_mm_csr.F:404-525
    def _init_stack_map(self):
        nm, nn, nk, nstacks = self.nm, self.nn, self.nk, self.nstacks
        self.m_map, mc_m, self.max_m = map_most_common(self.m_sizes, nm)
        self.n_map, mc_n, self.max_n = map_most_common(self.n_sizes, nn)
        self.k_map, mc_k, self.max_k = map_most_common(self.k_sizes, nk)
        descr = [None] * (nstacks + 1)  # 1-based
        smap = {}
        for m_map in range(1, nm + 2):
            m_size = mc_m[m_map - 1] if m_map <= nm else 777
            for k_map in range(1, nk + 2):
                k_size = mc_k[k_map - 1] if k_map <= nk else 888
                for n_map in range(1, nn + 2):
                    n_size = mc_n[n_map - 1] if n_map <= nn else 999
                    if m_map <= nm and k_map <= nk and n_map <= nn:
                        ps_g = (m_map - 1) * nn * nk + (k_map - 1) * nn + n_map
                        ps_g = nstacks - ps_g
                        smap[(n_map, k_map, m_map)] = ps_g
                        descr[ps_g] = dict(m=m_size, n=n_size, k=k_size, max_m=m_size, max_n=n_size, max_k=k_size,
                                           defined_mnk=True)
                    else:
                        smap[(n_map, k_map, m_map)] = nstacks
                        descr[nstacks] = dict(m=0, n=0, k=0, max_m=self.max_m, max_n=self.max_n, max_k=self.max_k,
                                              defined_mnk=False)
        flop_list = [-2 * descr[i]['m'] * descr[i]['n'] * descr[i]['k'] for i in range(1, nstacks)]
        flop_index = [p + 1 for p in stable_sort_index(flop_list)]  # 1-based old stack ids in new order
        new_descr = [None] * (nstacks + 1)
        for istack in range(1, nstacks):
            new_descr[istack] = descr[flop_index[istack - 1]]
        new_descr[nstacks] = descr[nstacks]
        pos = {old: new + 1 for new, old in enumerate(flop_index)}
        for key, old in smap.items():
            if old in pos:
                smap[key] = pos[old]
        self.stack_map, self.stacks_descr = smap, new_descr

    # src/mm/dbcsr_mm_accdrv.F:364-423
    def _stack_sort(self, params):
        order = stable_sort_index([p[5] for p in params])
        return [params[i][3:6] for i in order]

    def _stack_binning(self, params):
        nbins, binsize = self.nbins, self.binsize
        bins = [[] for _ in range(nbins)]
        out = []
        for p in params:
            val = p[3:6]
            # the reference multiplies in default (32-bit) INTEGER before widening: INT(val(3)*(val(3)+3), KIND=int_8) -- the
            # product wraps for c_first > 46339 (gfortran: two's complement); MODULO then floors, so the bin id stays in [0, nbins)
            c = int(val[2])
            prod = ((c * (c + 3)) + 2 ** 31) % 2 ** 32 - 2 ** 31
            bin_id = prod % nbins
            if len(bins[bin_id]) >= binsize:
                out.extend(bins[bin_id])
                bins[bin_id] = []
            bins[bin_id].append(val)
        for b in bins:
            out.extend(b)
        return out

    # src/mm/dbcsr_mm_sched.F:266 -> src/mm/dbcsr_mm_accdrv.F:433-541 (index work only)
    def _process(self, istack):
        params = self.stacks[istack - 1]
        d = self.stacks_descr[istack]
        flop_per_entry = 2 * d['max_m'] * d['max_n'] * d['max_k']
        if self.cfg_sort:
            dev = self._stack_sort(params) if flop_per_entry > self.min_flop_sort else self._stack_binning(params)
        else:
            dev = [p[3:6] for p in params]
        self.dispatched.append(dict(m=d['m'], n=d['n'], k=d['k'], max_m=d['max_m'], max_n=d['max_n'], max_k=d['max_k'],
                                    defined_mnk=d['defined_mnk'], stack_id=istack,
                                    host=np.array(params, dtype=np.int32).reshape(-1, 7),
                                    dev=np.array(dev, dtype=np.int32).reshape(-1, 3)))

    # src/mm/dbcsr_mm_csr.F:704-739
    def flush_stacks(self, purge=False):
        min_fill = 0 if purge else self.mm_stack_size * 3 // 4
        for i in range(1, self.nstacks + 1):
            if len(self.stacks[i - 1]) > min_fill:
                self._process(i)
                self.stacks[i - 1] = []

    # src/mm/dbcsr_mm_csr.F:178-359
    def csr_multiply_low(self, mi, mf, ki, kf, ai, af, bi, bf, a_index, b_index):
        a_row_p, a_blk_info = build_csr_index(mi, mf, ai, af, a_index)
        b_row_p, b_blk_info = build_csr_index(ki, kf, bi, bf, b_index)
        for a_row_l in range(mi, mf + 1):
            m_size = self.m_sizes[a_row_l - 1]
            mapped_row_size = self.m_map[m_size]
            use_eps = self.row_eps is not None and self.a_norms is not None
            a_row_eps = np.float32(self.row_eps[a_row_l - 1]) if use_eps else None
            for a_blk in range(a_row_p[a_row_l - mi] + 1, a_row_p[a_row_l - mi + 1] + 1):
                a_col_l, a_first, a_pos = a_blk_info[a_blk - 1]
                k_size = self.k_sizes[a_col_l - 1]
                mapped_k_size = self.k_map[k_size]
                for b_blk in range(b_row_p[a_col_l - ki] + 1, b_row_p[a_col_l - ki + 1] + 1):
                    b_col_l, b_first, b_pos = b_blk_info[b_blk - 1]
                    if use_eps:  # src/mm/dbcsr_mm_csr.F:270-278, single precision
                        if np.float32(np.float32(self.a_norms[a_pos - 1]) * np.float32(self.b_norms[b_pos - 1])) < a_row_eps:
                            self.skipped += 1
                            continue
                    if self.c_has_symmetry:  # "Don't calculate symmetric blocks", src/mm/dbcsr_mm_csr.F:280-292
                        c_row_logical = a_row_l if self.c_local_rows is None else int(self.c_local_rows[a_row_l - 1])
                        c_col_logical = b_col_l if self.c_local_cols is None else int(self.c_local_cols[b_col_l - 1])
                        if c_row_logical != c_col_logical and checker_tr(c_row_logical, c_col_logical):
                            continue
                    c_blk_id = self.c_hash.get((a_row_l, b_col_l), 0)
                    n_size = self.n_sizes[b_col_l - 1]
                    c_nze = m_size * n_size
                    if c_blk_id > 0:
                        offset = self.c_blk_p[c_blk_id - 1]
                    else:
                        if self.keep_sparsity:
                            continue
                        offset = self.datasize + 1
                        self.datasize += c_nze
                        self.c_row_i.append(a_row_l)
                        self.c_col_i.append(b_col_l)
                        self.c_blk_p.append(offset)
                        c_blk_id = len(self.c_blk_p)
                        self.c_hash[(a_row_l, b_col_l)] = c_blk_id
                    if c_nze == 0 or k_size == 0:
                        continue  # zero-sized block: C block created, nothing to multiply (the reference lets BLAS no-op it)
                    mapped_col_size = self.n_map[n_size]
                    ws = self.stack_map[(mapped_col_size, mapped_k_size, mapped_row_size)]
                    self.stacks[ws - 1].append((m_size, n_size, k_size, a_first, b_first, offset, c_blk_id))
                    self.flop += 2 * c_nze * k_size
                    if len(self.stacks[ws - 1]) >= self.mm_stack_size:
                        self.flush_stacks()

    # src/mm/dbcsr_mm_multrec.F:487-576
    def sparse_multrec(self, mi, mf, ni, nf, ki, kf, ai, af, a_index, bi, bf, b_index):
        if af < ai or bf < bi or mf < mi or nf < ni or kf < ki:
            return
        if af - ai + 1 <= self.multrec_limit and bf - bi + 1 <= self.multrec_limit:
            if af - ai + 1 > 0 and bf - bi + 1 > 0:
                self.csr_multiply_low(mi, mf, ki, kf, ai, af, bi, bf, a_index, b_index)
            return
        M, N, K = mf - mi + 1, nf - ni + 1, kf - ki + 1
        cut = 0
        if M >= max(N, K):
            cut = 1
        if K >= max(N, M):
            cut = 2
        if N >= max(M, K):
            cut = 3
        if cut == 1:
            s1 = M // 2
            acut = find_cut(a_index, ai, af, 0, mi + s1 - 1)
            self.sparse_multrec(mi, mi + s1 - 1, ni, nf, ki, kf, ai, acut - 1, a_index, bi, bf, b_index)
            self.sparse_multrec(mi + s1, mf, ni, nf, ki, kf, acut, af, a_index, bi, bf, b_index)
        elif cut == 2:
            s1 = K // 2
            acut = find_cut(a_index, ai, af, 1, ki + s1 - 1)
            bcut = find_cut(b_index, bi, bf, 0, ki + s1 - 1)
            self.sparse_multrec(mi, mf, ni, nf, ki, ki + s1 - 1, ai, acut - 1, a_index, bi, bcut - 1, b_index)
            self.sparse_multrec(mi, mf, ni, nf, ki + s1, kf, acut, af, a_index, bcut, bf, b_index)
        else:
            s1 = N // 2
            bcut = find_cut(b_index, bi, bf, 1, ni + s1 - 1)
            self.sparse_multrec(mi, mf, ni, ni + s1 - 1, ki, kf, ai, af, a_index, bi, bcut - 1, b_index)
            self.sparse_multrec(mi, mf, ni + s1, nf, ki, kf, ai, af, a_index, bcut, bf, b_index)

    # src/mm/dbcsr_mm_multrec.F:263-324 (+ setup_rec_index_2d, src/mm/dbcsr_mm_cannon.F:2910-2967)
    def multiply(self, a_list, b_list, a_norms=None, b_norms=None, row_eps=None):
        """a_norms/b_norms (aligned with a_list/b_list) + row_eps (per C row) switch the on-the-fly filter on."""
        nrow, ncol, nk = len(self.m_sizes), len(self.n_sizes), len(self.k_sizes)
        if a_norms is not None:
            # carry the norm with its block through the sort (the reference computes norms on the sorted images)
            a_list = [tuple(e[:3]) + (float(a_norms[i]),) for i, e in enumerate(a_list)]
            b_list = [tuple(e[:3]) + (float(b_norms[i]),) for i, e in enumerate(b_list)]
        a_index = rec_sort_index(1, nrow, 1, nk, list(a_list)) if len(a_list) > 1 else list(a_list)
        b_index = rec_sort_index(1, nk, 1, ncol, list(b_list)) if len(b_list) > 1 else list(b_list)
        if a_norms is not None:
            self.a_norms = np.array([e[3] for e in a_index], dtype=np.float32)
            self.b_norms = np.array([e[3] for e in b_index], dtype=np.float32)
            self.row_eps = None if row_eps is None else np.asarray(row_eps, dtype=np.float32)
            a_index = [e[:3] for e in a_index]
            b_index = [e[:3] for e in b_index]
        else:
            self.a_norms = self.b_norms = self.row_eps = None
        self.a_index, self.b_index = a_index, b_index
        self.sparse_multrec(1, nrow, 1, ncol, 1, nk, 1, len(a_index), a_index, 1, len(b_index), b_index)
        self.flush_stacks(purge=True)
        return self.dispatched
