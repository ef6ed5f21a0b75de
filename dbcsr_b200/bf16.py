"""
dbcsr_b200/bf16.py -- host side of the tiled BF16 SpGEMM (dbcsr_b200/csrc/smm_bf16_tiled.cuh, BASELINE.json config 4).

Extension of the DBCSR interface (DBCSR has no 16-bit type): C(FP32) = A * B with both operands rounded to BF16 once per panel.
For dense-ish products the multiply is driven by the block index (presence maps) instead of parameter stacks; everything that
computes runs on the device through the C ABI (libsmm_acc_b200_pack_bf16_rk, libsmm_acc_b200_bf16_spgemm) -- there is no CPU path.
"""
import numpy as np


class Bf16SpGemm:
    """C = A * B for panels whose block rows all have m, block columns n and k blocks k elements (m, n, k <= 32).

    A, B: dbcsr_b200.workload.Panel (BCSR-ordered block lists + FP64 data areas of column-major blocks).  The result is the
    dense block grid in BCSR order: block (rb, cb) at element offset (rb * ncb + cb) * m * n of an FP32 device buffer."""

    def __init__(self, acc, A, B, stream):
        self.acc, self.s = acc, stream
        sizes = [np.unique(x) for x in (A.row_sizes, B.col_sizes, A.col_sizes)]
        if any(u.size != 1 for u in sizes) or not np.array_equal(A.col_sizes, B.row_sizes):
            raise ValueError("the tiled BF16 kernel needs one block size per dimension")
        self.m, self.n, self.k = (int(u[0]) for u in sizes)
        if max(self.m, self.n, self.k) > 32:
            raise ValueError("block dimensions above 32 are not supported by the BF16 kernels")
        self.nrb, self.ncb, self.nkb = int(A.row_sizes.size), int(B.col_sizes.size), int(A.col_sizes.size)
        m, n, k = self.m, self.n, self.k
        # ---- FP64 panels -> BF16 operand tiles (one per block, block order = tile order); B is packed straight from its
        #      untransposed k x n column-major blocks: operand row = block column index, element (col, kk) at src[kk + col*k]
        #      A tiles are stored in block-COLUMN order (slot = rank in (k block, block row) order), B tiles in block-row order at the
        #      2 KB slot pitch: tiles of blocks that are adjacent in a k block's operand are then adjacent in memory and the kernel
        #      fetches a run of them with one bulk copy
        pa, pb = acc.bf16_rk_slot_bytes(m, False), acc.bf16_rk_slot_bytes(n, True)
        self.a_tiles = acc.dev_alloc(max(A.nblks, 1) * pa)
        self.b_tiles = acc.dev_alloc(max(B.nblks, 1) * pb)
        a_slot = np.empty(A.nblks, dtype=np.int32)
        a_slot[np.lexsort((A.rows, A.cols))] = np.arange(A.nblks, dtype=np.int32)
        d_slot = acc.to_device(a_slot, stream) if A.nblks else None
        self.h2d_bytes = a_slot.nbytes
        for panel, tiles, rows, rs, ks, pitch, slot in ((A, self.a_tiles, m, 1, m, pa, d_slot), (B, self.b_tiles, n, k, 1, pb, None)):
            if panel.nblks:
                d = acc.to_device(panel.data, stream)
                self.h2d_bytes += panel.data.nbytes
                acc.pack_bf16_rk(d.ptr, panel.nblks, rows, k, rs, ks, tiles.ptr, pitch, slot.ptr if slot is not None else None, stream)
                acc.stream_sync(stream)
                d.free()
        if d_slot is not None:
            d_slot.free()
        # ---- presence maps: slot of the tile of block (rb, kb) of A / (kb, cb) of B, or -1
        a_map = np.full((self.nkb, self.nrb), -1, dtype=np.int32)
        a_map[A.cols - 1, A.rows - 1] = a_slot
        b_map = np.full((self.nkb, self.ncb), -1, dtype=np.int32)
        b_map[B.rows - 1, B.cols - 1] = np.arange(B.nblks, dtype=np.int32)
        c_elems = self.nrb * self.ncb * m * n
        if c_elems >= 2 ** 31:
            raise ValueError("C exceeds int32 element offsets")
        c_off = (np.arange(self.nrb * self.ncb, dtype=np.int64) * (m * n)).astype(np.int32)
        self.d_a_map, self.d_b_map, self.d_c_off = acc.to_device(a_map, stream), acc.to_device(b_map, stream), acc.to_device(c_off, stream)
        self.h2d_bytes += a_map.nbytes + b_map.nbytes + c_off.nbytes
        self.c_elems = c_elems
        self.d_c = acc.dev_alloc(4 * max(c_elems, 1))
        # work actually requested: block products = sum_k (#A blocks in block column k) * (#B blocks in block row k)
        na_k = np.bincount(A.cols - 1, minlength=self.nkb).astype(np.int64)
        nb_k = np.bincount(B.rows - 1, minlength=self.nkb).astype(np.int64)
        self.products = int(np.dot(na_k, nb_k))
        self.flop = 2 * m * n * k * self.products
        # issued (padded) tensor-core work: per tile and k block with at least one A block, one M=128 x N=32 x K=32 MMA pair per
        # existing B block of the tile's 16 block columns
        bpt = min(16 // ((m + 7) // 8), 5)
        a_any = np.zeros((self.nkb, (self.nrb + bpt - 1) // bpt), dtype=np.int64)
        np.maximum.at(a_any, (A.cols - 1, (A.rows - 1) // bpt), 1)
        nb = 15 if (acc.get_tunable("bf16_a_tmem") and not acc.get_tunable("bf16_plan")) else 16  # block columns per tile (smm_bf16_tiled.cuh)
        b_cnt = np.zeros((self.nkb, (self.ncb + nb - 1) // nb), dtype=np.int64)
        np.add.at(b_cnt, (B.rows - 1, (B.cols - 1) // nb), 1)
        self.mma_pairs = int(np.einsum("kr,kc->", a_any, b_cnt))
        self.issued_flop = self.mma_pairs * 2 * 128 * 32 * 32

    def run(self):
        self.acc.bf16_spgemm(self.a_tiles.ptr, self.d_a_map.ptr, self.b_tiles.ptr, self.d_b_map.ptr, self.d_c.ptr, self.d_c_off.ptr, self.nrb,
                             self.ncb, self.nkb, self.m, self.n, self.k, self.s)

    def result(self):
        """FP32 C as a (nrb, ncb, n, m) array: [rb, cb] is the column-major m x n block (index [rb, cb, col, row])."""
        c = self.acc.to_host(self.d_c, (max(self.c_elems, 1),), np.float32, self.s)
        return c[:self.c_elems].reshape(self.nrb, self.ncb, self.n, self.m)

    def close(self):
        for d in (self.a_tiles, self.b_tiles, self.d_a_map, self.d_b_map, self.d_c_off, self.d_c):
            d.free()
