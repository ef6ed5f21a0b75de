// dbcsr_b200/csrc/smm_bf16_tiled.cuh -- BF16 block-sparse GEMM as a TILED SpGEMM on tcgen05 (BASELINE.json config 4: 23x23 blocks,
// 50 % occupation, BF16 operands, FP32 accumulate).  Extension of the DBCSR ABI like smm_bf16.cuh (DBCSR has no 16-bit type).
//
// Why not the stack: a C-sorted parameter stack hands the kernel one 23x23x23 product at a time; each product then moves two
// operand tiles (2 x 1.1 KB) for 24 kflop and fills 18 % of a 128-row MMA -- the per-entry kernel (smm_bf16.cuh) is bound by
// L2 -> shared-memory traffic at 2.4 % of the tensor peak.  At 50 % occupation the product is dense-ish, so this kernel is driven
// by the BLOCK INDEX instead (presence maps of A and B), like a tiled GEMM:
//   * C tile = BPT block rows (BPT = 16 / ceil(m/8) = 5 for 23-row blocks: 120 of the 128 MMA rows) x 16 block columns; its FP32
//     accumulators live in TMEM for the whole k loop (128 lanes x 16 x 32 columns = all 512 columns) and are written to C ONCE by
//     plain coalesced stores -- no atomics, no memset of C, no read of C;
//   * per k block: the A blocks of the BPT rows that exist are staged by one TMA bulk copy each (1.5 KB, operand layout, see
//     below) into their row slot of the 128 x 32 A operand (absent slots are zero-filled once and stay zero until reused), the B
//     blocks that exist among the 16 columns into their column slot (2 KB pitch = 32 operand rows); ONE elected thread then issues
//     two tcgen05.mma (M=128, K=16) per RUN of adjacent existing B blocks, N = 32 x run length, into the accumulators of those
//     block columns: absent B blocks cost nothing, absent A blocks cost padding rows (the MMA is an outer product: it cannot
//     skip rows per k).  Merging runs halves the MMA count (4.3 instead of 8 per k step at 50 % occupation) and with it the
//     re-reads of the A operand from shared memory, which bound small-N MMAs;
//   * the A operand of a k block is shared by up to 16 MMAs pairs, a B block by 5 block rows: 15.8 KB of L2 traffic per k block
//     and tile instead of 8 x 5 x 2.2 KB = 90 KB for the same products through the stack kernel;
//   * warp-specialised, persistent (one CTA per SM, static round-robin over tiles ordered so that the ~148 concurrently running
//     tiles form a ~12 x 12 patch of the tile grid and share their A/B panels in L2): warp 0 = TMA producer of the B blocks,
//     warp 6 = TMA producer of the A blocks (+ zero-fill of absent A slots), warp 1 = MMA issuer (one elected lane), warps 2-5 =
//     epilogue (tcgen05.ld -> st.global).  The first version had ONE producer warp and one bulk copy per block: ncu showed that
//     warp executing 259 instructions per k block, 1230 cycles, with the MMA warp waiting for it 40 % of its time -- so the
//     copies of ADJACENT existing blocks are merged (tiles of adjacent blocks are adjacent in memory: A tiles are packed in
//     block-column order, B tiles in block-row order at the 2 KB slot pitch) and the work is split over two warps.
//
// Operand tile format ("rk", one per block, ROWS <= 32, K padded to 32): element (row, kk) at byte
//     (row / 8) * 512 + (kk / 8) * 128 + (row % 8) * 16 + (kk % 8) * 2
// i.e. K-major no-swizzle core matrices (8 rows x 8 k = 128 B), the 4 k groups of a row group contiguous: a block is ONE
// contiguous piece (ceil(ROWS/8) * 512 B) and consecutive blocks stack along M/N with a uniform row-group stride -- the
// shared-memory descriptor is LBO (k-group stride) = 128 B, SBO (row-group stride) = 512 B.  Padding rows / k are zeros.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "smm_bf16.cuh"

namespace smm {

constexpr int BT_STAGES = 5;
constexpr int BT_NB = 16;            // block columns per tile (TMEM: 16 x 32 columns)
constexpr int BT_THREADS = 224;      // B producer warp, MMA warp, 4 epilogue warps, A producer warp
constexpr int BT_A_BYTES = 16 * 512; // A operand of one k block: 16 row groups x 4 k groups x 128 B
constexpr int BT_KC = 8;             // presence-map entries prefetched per lane
constexpr int BT_B_SLOT = 2048;      // B slot pitch: 32 operand rows (4 row groups), so that adjacent slots form one N = 32 r operand
constexpr int BT_FLAG_MERGE_RUNS = 1;
constexpr int BT_FLAG_A_TMEM = 2;     // stage the A operand of every k block in TMEM (tcgen05.cp) and multiply from there
constexpr int BT_NB_A_TMEM = 15;      // block columns per tile in that mode: 15 x 32 accumulator columns + 2 x 16 columns of A

struct BtGeom {
  int rg_a, rg_b;      // row groups per A / B block
  int bpt;             // A blocks per tile (16 / rg_a)
  int tile_a, tile_b;  // bytes per packed block
  int stage;           // bytes per pipeline stage
};
__host__ __device__ inline BtGeom bt_geom(int m, int n) {
  BtGeom g;
  g.rg_a = (m + 7) / 8;
  g.rg_b = (n + 7) / 8;
  g.bpt = 16 / g.rg_a;
  if (g.bpt > 5) g.bpt = 5;  // the producer warp has five A lanes; small blocks (m <= 16) leave MMA rows unused
  g.tile_a = g.rg_a * 512;
  g.tile_b = g.rg_b * 512;
  // B slots at a pitch of 32 rows: the row groups a block does not fill stay zero (never written after the initial clear)
  g.stage = BT_A_BYTES + BT_NB * BT_B_SLOT;
  return g;
}
__host__ __device__ inline size_t bt_smem_bytes(const BtGeom& g) { return 1024 + (size_t)BT_STAGES * g.stage; }

// FP64 block (element (row,kk) at src[row*row_stride + kk*k_stride]) -> BF16 "rk" tile, round to nearest even; one warp per block.
// Tile of block b goes to dst + slot * pitch, slot = dst_slot[b] (nullptr: b); pitch >= tile size, the gap is zero-filled.
__global__ void pack_bf16_rk_kernel(const double* __restrict__ src, int nblks, int rows, int kdim, int row_stride, int k_stride,
                                    unsigned char* __restrict__ dst, int pitch, const int* __restrict__ dst_slot) {
  const int wpc = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = (rows + 7) / 8;
  const int nelem = pitch / 2;  // bf16 elements per slot incl. padding
  for (int b = blockIdx.x * wpc + warp; b < nblks; b += gridDim.x * wpc) {
    const double* __restrict__ s = src + (size_t)b * rows * kdim;
    const int slot = dst_slot != nullptr ? __ldg(dst_slot + b) : b;
    unsigned short* __restrict__ d = reinterpret_cast<unsigned short*>(dst + (size_t)slot * pitch);
    for (int i = lane; i < nelem; i += 32) {
      const int kk8 = i & 7, r8 = (i >> 3) & 7, kgi = (i >> 6) & 3, rgi = i >> 8;
      const int row = rgi * 8 + r8, kk = kgi * 8 + kk8;
      float v = 0.f;
      if (rgi < rg && row < rows && kk < kdim) v = (float)s[(size_t)row * row_stride + (size_t)kk * k_stride];
      unsigned int u = __float_as_uint(v);
      u += 0x7fffu + ((u >> 16) & 1u);
      d[i] = (unsigned short)(u >> 16);
    }
  }
}

// one lane of a converged warp (elect.sync): the predicate is warp-uniform knowledge for the compiler, unlike `lane == 0`
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "elect.sync _|p, 0xffffffff;\n"
    "selp.u32 %0, 1, 0, p;\n"
    "}\n"
    : "=r"(pred));
  return pred;
}

// tcgen05.cp: 128 lanes x 256 bit from the shared-memory matrix `desc` (same descriptor format as an MMA operand: 16 row groups
// at SBO, two 16-byte k chunks at LBO) into TMEM columns [taddr, taddr + 8)
__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
// tcgen05.mma with the A operand in TMEM (M = 128 rows = lanes, K = 16 BF16 = 8 columns), B from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "setp.ne.b32 p, %4, 0;\n"
    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
    "}\n" ::"r"(tmem_d),
    "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
    : "memory");
}

// tile t -> (row group, column group): panels of BT_PANEL row groups, inside a panel column groups outer, row groups inner, so
// that gridDim.x consecutive tiles cover about 12 x 12 tiles
constexpr int BT_PANEL = 12;
__device__ __forceinline__ void bt_tile_coords(int t, int n_rg, int n_cg, int& rg, int& cg) {
  const int per_panel = BT_PANEL * n_cg;
  const int p = t / per_panel, r = t - p * per_panel;
  const int rows_here = min(BT_PANEL, n_rg - p * BT_PANEL);
  cg = r / rows_here;
  rg = p * BT_PANEL + (r - cg * rows_here);
}
__host__ __device__ inline int bt_num_tiles(int n_rg, int n_cg) { return n_rg * n_cg; }

// a_map[kb * nrb + rb] / b_map[kb * ncb + cb]: index of the packed tile of block (rb, kb) of A / (kb, cb) of B, or -1;
// c_off[rb * ncb + cb]: 0-based element offset of C block (rb, cb) (column-major m x n, FP32) or -1 = not stored.
__global__ void __launch_bounds__(BT_THREADS, 1)
  smm_bf16_tiled_kernel(const unsigned char* __restrict__ a_tiles, const int* __restrict__ a_map, const unsigned char* __restrict__ b_tiles,
                        const int* __restrict__ b_map, float* __restrict__ c_data, const int* __restrict__ c_off, int nrb, int ncb, int nkb,
                        int m, int n, int flags) {
  // Block columns per tile.  SS mode: 16 (all 512 TMEM columns are accumulators).  A-in-TMEM mode: 15, the last 32 columns hold
  // the A operand of the current / next k block: small-N MMAs are bound by re-reading the 128-row A operand from shared memory
  // for every instruction (ncu: 47 operand wavefronts per M=128 x N=60 MMA, 32 of them A); one tcgen05.cp per 16 k stages it in
  // TMEM once and every MMA of the stage reads it from there.
  const int nb = (flags & BT_FLAG_A_TMEM) ? BT_NB_A_TMEM : BT_NB;
  extern __shared__ __align__(1024) unsigned char bt_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const BtGeom g = bt_geom(m, n);
  const int n_rg = (nrb + g.bpt - 1) / g.bpt, n_cg = (ncb + nb - 1) / nb;
  const int n_tiles = bt_num_tiles(n_rg, n_cg);

  uint64_t* full = reinterpret_cast<uint64_t*>(bt_smem);  // [BT_STAGES]
  uint64_t* empty = full + BT_STAGES;                     // [BT_STAGES]
  uint64_t* tmem_full = empty + BT_STAGES;                // [1]
  uint64_t* tmem_empty = tmem_full + 1;                   // [1]
  uint32_t* meta = reinterpret_cast<uint32_t*>(tmem_empty + 1);  // [BT_STAGES] block-column mask of the stage
  uint32_t* tile_inited = meta + BT_STAGES;                      // [1] block columns of the finished tile that received an MMA
  uint32_t* tmem_ptr = tile_inited + 1;
  unsigned char* stages = bt_smem + 1024;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");  // pack kernels / map uploads of this stream are complete and visible

  // all operand bytes start as zeros: padding row group 15 of the A operand, the slack behind the B slots, never-loaded slots
  for (size_t i = (size_t)threadIdx.x * 16; i < (size_t)BT_STAGES * g.stage; i += (size_t)BT_THREADS * 16)
    *reinterpret_cast<uint4*>(stages + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < BT_STAGES; ++s) {
      mbar_init(&full[s], 2);  // the two producer warps
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic zero-fill -> visible to the TMA / tensor-core proxies
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 || warp == 6) {
    // ===================================== TMA producers (whole warps) =====================================
    // Both warps read the same presence-map entries (lane l < bpt: A slot l, lane 5 + c: B slot c) and derive identical masks;
    // warp 6 copies the A blocks and keeps absent A slots zero, warp 0 copies the B blocks and publishes the stage's column mask.
    const bool a_role = warp == 6;
    const int bpt = g.bpt;
    const bool is_a = lane < bpt, is_b = lane >= 5 && lane < 5 + nb;
    const bool mine = a_role ? is_a : is_b;
    uint32_t zero_state = 0;  // bit (5 * stage + slot): A slot is known to hold zeros (everything is zero at start)
    for (int s0 = 0; s0 < BT_STAGES; ++s0) zero_state |= 0x1fu << (5 * s0);
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      int rg, cg;
      bt_tile_coords(t, n_rg, n_cg, rg, cg);
      // this lane's presence-map column: &map[0 * stride + col]; valid = inside the matrix
      const int rb = rg * bpt + lane, cb = cg * nb + (lane - 5);
      const bool valid = is_a ? (rb < nrb) : (is_b ? (cb < ncb) : false);
      const int* __restrict__ mp = is_a ? (a_map + rb) : (b_map + cb);
      const int mstride = is_a ? nrb : ncb;
      const unsigned char* __restrict__ tiles = is_a ? a_tiles : b_tiles;
      const uint32_t pitch = is_a ? (uint32_t)g.tile_a : (uint32_t)BT_B_SLOT;  // tile pitch in global AND shared memory
      const uint32_t slot_off = is_a ? (uint32_t)lane * pitch : (uint32_t)BT_A_BYTES + (uint32_t)(lane - 5) * pitch;
      int cur[BT_KC], nxt[BT_KC];
#pragma unroll
      for (int j = 0; j < BT_KC; ++j) cur[j] = (valid && j < nkb) ? __ldg(mp + (size_t)j * mstride) : -1;
      for (int k0 = 0; k0 < nkb; k0 += BT_KC) {
#pragma unroll
        for (int j = 0; j < BT_KC; ++j) nxt[j] = (valid && k0 + BT_KC + j < nkb) ? __ldg(mp + (size_t)(k0 + BT_KC + j) * mstride) : -1;
#pragma unroll
        for (int j = 0; j < BT_KC; ++j) {
          if (k0 + j < nkb) {  // warp-uniform
            const int idx = cur[j];
            const unsigned mask = __ballot_sync(0xffffffffu, idx >= 0);
            const uint32_t am = mask & 0x1fu;
            const uint32_t bm = am != 0 ? ((mask >> 5) & 0xffffu) : 0u;  // no A block in these rows: nothing to multiply
            // a lane CONTINUES the copy of its left neighbour when both blocks exist and their tiles are adjacent in memory
            // (never across the A | B boundary): one bulk copy per run instead of one per block
            const int idx_left = __shfl_up_sync(0xffffffffu, idx, 1);
            const bool cont = idx >= 0 && idx_left >= 0 && idx == idx_left + 1 && lane != 0 && lane != 5;
            const unsigned contm = __ballot_sync(0xffffffffu, cont);
            mbar_wait(&empty[s], ph ^ 1u);
            unsigned char* stg = stages + (size_t)s * g.stage;
            if (a_role && bm != 0) {
              // absent A slots must read as zeros: fill those that held data
              const uint32_t zs = (zero_state >> (5 * s)) & 0x1fu;
              uint32_t fill = ~am & ~zs & ((1u << bpt) - 1u);
              if (fill != 0) {
                while (fill != 0) {
                  const int slot = __ffs(fill) - 1;
                  fill &= fill - 1;
                  unsigned char* dst = stg + (size_t)slot * g.tile_a;
                  for (int o = lane * 16; o < g.tile_a; o += 32 * 16) *reinterpret_cast<uint4*>(dst + o) = make_uint4(0, 0, 0, 0);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              }
              zero_state = (zero_state & ~(0x1fu << (5 * s))) | ((~am & 0x1fu) << (5 * s));
            }
            if (lane == 0) {
              if (!a_role) meta[s] = bm;
              const uint32_t bytes = bm == 0 ? 0u : (a_role ? (uint32_t)__popc(am) * (uint32_t)g.tile_a : (uint32_t)__popc(bm) * (uint32_t)BT_B_SLOT);
              mbar_expect_tx(&full[s], bytes);  // this warp's arrival; the phase completes when both warps' bytes have landed
            }
            __syncwarp();
            if (bm != 0 && mine && idx >= 0 && !cont) {
              const unsigned above = contm >> (lane + 1);  // lanes <= 20
              const uint32_t len = (uint32_t)__ffs(~above);  // 1 + number of continuing lanes to the right
              bulk_g2s(stg + slot_off, tiles + (size_t)idx * pitch, len * pitch, &full[s]);
            }
            if (++s == BT_STAGES) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < BT_KC; ++j) cur[j] = nxt[j];
      }
    }
  }
  else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // The whole warp walks the loop (barrier waits and masks are warp-uniform); ONE elected lane issues the MMAs and commits, so
    // that descriptors live in uniform registers and no per-lane serialisation is generated around tcgen05.mma.
    // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format BF16 (1) [7,10),[10,13), K-major A and B,
    // N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const bool merge = (flags & BT_FLAG_MERGE_RUNS) != 0, a_tmem = (flags & BT_FLAG_A_TMEM) != 0;
    uint32_t it = 0, tile_no = 0, ph = 0;
    int s = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_no) {
      mbar_wait(tmem_empty, (tile_no & 1u) ^ 1u);  // the epilogue has drained the previous tile's accumulators
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t inited = 0;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        mbar_wait(&full[s], ph);
        const uint32_t bm = meta[s];
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t sa = smem_u32(stages + (size_t)s * g.stage);
          const uint64_t adesc0 = umma_desc(sa, 128u, 512u), adesc1 = umma_desc(sa + 256u, 128u, 512u);
          const uint64_t bdesc0 = umma_desc(sa + (uint32_t)BT_A_BYTES, 128u, 512u);
          uint32_t todo = bm;
          if (a_tmem) {
            // A operand: shared memory -> TMEM once per stage (2 x 128 lanes x 256 bit = k 0..15 | 16..31), double buffered; the
            // copies and the MMAs execute in issue order, so no barrier is needed between them
            const uint32_t ta = tmem_base + (uint32_t)(BT_NB_A_TMEM * 32) + 16u * (it & 1u);
            if (bm != 0) {
              utccp_128x256b(ta, adesc0);
              utccp_128x256b(ta + 8u, adesc1);
            }
            while (todo != 0) {
              const int c0 = __ffs(todo) - 1;
              int r = 1;
              if (merge) {
                r = __ffs(~(todo >> c0)) - 1;
                const uint32_t st = inited >> c0;
                const int same = (st & 1u) ? (__ffs(~st) - 1) : (st == 0 ? 32 : __ffs(st) - 1);
                r = min(min(r, same), 8);
              }
              const uint32_t idesc = idesc_base | ((uint32_t)(4 * r) << 17);
              const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)c0 * (uint32_t)(BT_B_SLOT >> 4));
              const uint32_t d = tmem_base + 32u * (uint32_t)c0;
              umma_bf16_ts(d, ta, bd, idesc, (inited >> c0) & 1u);
              umma_bf16_ts(d, ta + 8u, bd + 16u, idesc, 1u);
              todo &= ~(((1u << r) - 1u) << c0);
            }
          }
          else if (merge && (inited & bm) == bm) {
            // steady state: every touched accumulator is initialised -> runs of adjacent existing blocks, always accumulating
            while (todo != 0) {
              const int c0 = __ffs(todo) - 1;
              const int r = min(__ffs(~(todo >> c0)) - 1, 8);                                       // N <= 256
              const uint32_t idesc = idesc_base | ((uint32_t)(4 * r) << 17);                        // N = 32 r
              const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)c0 * (uint32_t)(BT_B_SLOT >> 4));  // start address in 16-byte units
              const uint32_t d = tmem_base + 32u * (uint32_t)c0;
              umma_bf16(d, adesc0, bd, idesc, 1u);
              umma_bf16(d, adesc1, bd + 16u, idesc, 1u);  // k = 16..31: +256 B
              todo &= ~(((1u << r) - 1u) << c0);
            }
          }
          else {
            while (todo != 0) {
              const int c0 = __ffs(todo) - 1;
              int r = 1;
              if (merge) {
                // adjacent existing blocks whose accumulators are in the same state (all initialised or all fresh), at most 8
                r = __ffs(~(todo >> c0)) - 1;
                const uint32_t st = inited >> c0;
                const int same = (st & 1u) ? (__ffs(~st) - 1) : (st == 0 ? 32 : __ffs(st) - 1);
                r = min(min(r, same), 8);
              }
              const uint32_t idesc = idesc_base | ((uint32_t)(4 * r) << 17);
              const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)c0 * (uint32_t)(BT_B_SLOT >> 4));
              const uint32_t d = tmem_base + 32u * (uint32_t)c0;
              umma_bf16(d, adesc0, bd, idesc, (inited >> c0) & 1u);
              umma_bf16(d, adesc1, bd + 16u, idesc, 1u);
              todo &= ~(((1u << r) - 1u) << c0);
            }
          }
          umma_commit(&empty[s]);  // the stage may be refilled once these MMAs have read it
        }
        inited |= bm;
        __syncwarp();
        if (++s == BT_STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (elect_one()) {
        *tile_inited = inited;
        __threadfence_block();  // the epilogue reads tile_inited after the (asynchronous) commit-arrive on tmem_full
        umma_commit(tmem_full);
      }
      __syncwarp();
    }
  }
  else if (warp < 6) {
    // ===================================== epilogue (4 warps = 128 TMEM lanes) =====================================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;      // row of the 128-row tile
    const int rg_rows = g.rg_a * 8;     // row pitch of a block inside the tile
    const int blk = row / rg_rows, r_in = row - blk * rg_rows;
    uint32_t tile_no = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_no) {
      int rg, cg;
      bt_tile_coords(t, n_rg, n_cg, rg, cg);
      const int rb = rg * g.bpt + blk;
      const bool row_ok = blk < g.bpt && r_in < m && rb < nrb;
      mbar_wait(tmem_full, tile_no & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t inited = *tile_inited;
      for (int c = 0; c < nb; ++c) {
        const int cb = cg * nb + c;
        if (cb >= ncb) break;  // warp-uniform
        uint32_t r[32];
        asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
            "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
            "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
            "=r"(r[31])
          : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + 32u * (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row_ok) {
          const int off = __ldg(c_off + (size_t)rb * ncb + cb);
          if (off >= 0) {
            float* __restrict__ dst = c_data + (size_t)off + r_in;
            const bool have = (inited >> c) & 1u;
#pragma unroll
            for (int col = 0; col < 32; ++col)
              if (col < n) dst[(size_t)col * m] = have ? __uint_as_float(r[col]) : 0.f;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace smm
