// dbcsr_b200/csrc/smm_bf16_plan.cuh -- tiled BF16 SpGEMM, PLANNED variant (round 2, second step; smm_bf16_tiled.cuh has the design).
//
// ncu of the first tiled kernel showed all three helper warps instruction-latency bound: ~128 instructions per k block in each
// TMA producer (presence-map loads, ballots, shuffles, bit scans, zero-fill state) and ~158 in the MMA issuer (run detection,
// accumulator-state tracking), 6-8 cycles each for a lone warp => ~1000 cycles per k block against a tensor-pipe floor of ~340.
// Everything those warps derived from the presence maps depends only on (row group, k block) or (column group, k block) -- not on
// the tile -- so a small kernel (`bt_plan_kernel`) now derives it ONCE per multiply:
//   * A plan, per (row group, k block): up to 5 ready-made bulk-copy commands {source address, stage offset, bytes}; absent A
//     slots are filled by a copy from a zero tile in global memory (no zero-fill loop, no per-stage state);
//   * B plan, per (column group, k block): the copy commands of the runs of adjacent existing B blocks, and the MMA runs
//     (first block column, run length <= 8) packed into 8 bytes.
// The producer warps only load a command per lane and issue it; the MMA issuer walks the packed runs.  The accumulators are
// ALWAYS accumulated into: the epilogue warps zero the TMEM columns they have just read (tcgen05.st), so no warp tracks which
// block columns have been initialised.
#pragma once
#include "smm_bf16_tiled.cuh"

namespace smm {

// Third step: with the bookkeeping gone the kernel ran at 745 cycles per k block with 5 stages of 40 KB (8 KB of A + sixteen 2 KB B
// slots, of which half are used at 50 % occupation): the time of a stage's round trip (TMA latency under load + MMAs) divided by
// the stages in flight.  So the B blocks of a k block are now PACKED (only existing blocks occupy shared memory): a ring of 2 KB
// slots managed by the B producer (allocation in ring order, a stage that would cross the nominal end starts at slot 0, older
// stages are awaited oldest-first when their slots are needed), 8 stage entries instead of 5, ~1.6x the bytes in flight.
constexpr int BP_NS = 8;          // stage entries (A: fixed 8 KB slots; B: ring allocation)
constexpr int BP_RING = 81;       // physical B slots: (227 KB - 1 KB - 8 x 8 KB) / 2 KB
constexpr int BP_RING_NOM = 66;   // a stage starts below this slot (it may extend up to 15 slots further)
constexpr int BP_RA = 8;   // command slots per row of the A plan (<= 5 used)
constexpr int BP_RB = 16;  // command slots per row of the B plan
constexpr int BP_KC = 8;   // plan rows prefetched per lane

struct BtPlanPtrs {
  uint4* a_cmd;           // [n_rg][nkb][BP_RA]: x,y = source address, z = stage offset | bytes << 16, w (slot 0) = bytes of the row
  uint4* b_cmd;           // [n_cg][nkb][BP_RB]
  uint4* m_runs;          // [n_cg][nkb]: 8 x 16 bit (0x80 | (run - 1) << 4 | first block column | rank of the first block among the existing ones << 8)
  unsigned char* a_any;   // [n_rg][nkb]: the row group has an A block in this k block
  unsigned char* zeros;   // 5 A tiles of zeros
};
inline size_t bp_smem_bytes() { return 1024 + (size_t)BP_NS * BT_A_BYTES + (size_t)BP_RING * BT_B_SLOT; }
inline size_t bt_plan_bytes(int n_rg, int n_cg, int nkb, size_t* off /* [5] */) {
  size_t o = 0;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  off[0] = o, o += up((size_t)n_rg * nkb * BP_RA * sizeof(uint4));
  off[1] = o, o += up((size_t)n_cg * nkb * BP_RB * sizeof(uint4));
  off[2] = o, o += up((size_t)n_cg * nkb * sizeof(uint4));
  off[3] = o, o += up((size_t)n_rg * nkb);
  off[4] = o, o += 5 * 2048;
  return o;
}

__device__ __forceinline__ uint4 bt_cmd(const unsigned char* src, uint32_t dst, uint32_t bytes) {
  const unsigned long long a = (unsigned long long)src;
  return make_uint4((uint32_t)a, (uint32_t)(a >> 32), dst | (bytes << 16), 0u);
}

__global__ void bt_plan_kernel(const unsigned char* __restrict__ a_tiles, const int* __restrict__ a_map, const unsigned char* __restrict__ b_tiles,
                               const int* __restrict__ b_map, int nrb, int ncb, int nkb, int m, int n, int nb, BtPlanPtrs P) {
  const BtGeom g = bt_geom(m, n);
  const int n_rg = (nrb + g.bpt - 1) / g.bpt, n_cg = (ncb + nb - 1) / nb;
  const long long total = (long long)(n_rg + n_cg) * nkb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int G = (int)(i / nkb), kb = (int)(i - (long long)G * nkb);
    if (G < n_rg) {
      uint4* out = P.a_cmd + ((size_t)G * nkb + kb) * BP_RA;
      int idx[5];
      bool any = false;
#pragma unroll
      for (int l = 0; l < 5; ++l) {
        const int rb = G * g.bpt + l;
        idx[l] = (l < g.bpt && rb < nrb) ? a_map[(size_t)kb * nrb + rb] : -1;
        any = any || idx[l] >= 0;
      }
      P.a_any[(size_t)G * nkb + kb] = any ? 1 : 0;
      // an absent slot needs zeros only if the stage's A slot may hold data there: the stage entry was last written BP_NS k blocks
      // earlier in this tile (k blocks without any A block are skipped by the producer); at the start of a tile: unknown => all
      bool dirty[5];
      {
        int kp = kb - BP_NS;
        bool found = false;
        for (int tries = 0; tries < 4 && kp >= 0 && !found; ++tries) {
          bool anyp = false;
#pragma unroll
          for (int l = 0; l < 5; ++l) {
            const int rb = G * g.bpt + l;
            dirty[l] = (l < g.bpt && rb < nrb) ? a_map[(size_t)kp * nrb + rb] >= 0 : false;
            anyp = anyp || dirty[l];
          }
          if (anyp)
            found = true;
          else
            kp -= BP_NS;
        }
        if (!found) {
#pragma unroll
          for (int l = 0; l < 5; ++l) dirty[l] = true;
        }
      }
      int nc = 0;
      uint32_t total_b = 0;
      if (any) {
        int l = 0;
        while (l < g.bpt) {
          int len = 1;
          const unsigned char* src;
          if (idx[l] >= 0) {
            while (l + len < g.bpt && idx[l + len] == idx[l] + len) ++len;  // adjacent tiles are adjacent in memory
            src = a_tiles + (size_t)idx[l] * g.tile_a;
          }
          else {
            if (!dirty[l]) {  // still zero from an earlier fill
              ++l;
              continue;
            }
            while (l + len < g.bpt && idx[l + len] < 0 && dirty[l + len]) ++len;
            src = P.zeros;
          }
          const uint32_t bytes = (uint32_t)(len * g.tile_a);
          out[nc++] = bt_cmd(src, (uint32_t)(l * g.tile_a), bytes);
          total_b += bytes;
          l += len;
        }
      }
      for (int c = nc; c < BP_RA; ++c) out[c] = make_uint4(0, 0, 0, 0);
      out[0].w = total_b;
    }
    else {
      const int cg = G - n_rg;
      uint4* out = P.b_cmd + ((size_t)cg * nkb + kb) * BP_RB;
      uint32_t bm = 0;
      int first_idx[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int cb = cg * nb + c;
        first_idx[c] = (c < nb && cb < ncb) ? b_map[(size_t)kb * ncb + cb] : -1;
        if (first_idx[c] >= 0) bm |= 1u << c;
      }
      int nc = 0;
      uint32_t total_b = 0;
      uint32_t runs[4] = {0u, 0u, 0u, 0u};
      int nr = 0;
      int c = 0;
      while (c < nb) {  // copy commands: existing blocks whose tiles are adjacent in memory; destination = packed (rank * 2 KB)
        if (first_idx[c] < 0) {
          ++c;
          continue;
        }
        int len = 1;
        while (c + len < nb && first_idx[c + len] == first_idx[c] + len) ++len;
        const uint32_t bytes = (uint32_t)(len * BT_B_SLOT);
        const uint32_t rank = (uint32_t)__popc(bm & ((1u << c) - 1u));
        out[nc++] = bt_cmd(b_tiles + (size_t)first_idx[c] * BT_B_SLOT, rank * (uint32_t)BT_B_SLOT, bytes);
        total_b += bytes;
        c += len;
      }
      for (int q = nc; q < BP_RB; ++q) out[q] = make_uint4(0, 0, 0, 0);
      out[0].w = total_b;
      c = 0;
      while (c < nb) {  // MMA runs: adjacent existing blocks, N = 32 * run <= 256
        if (!((bm >> c) & 1u)) {
          ++c;
          continue;
        }
        int r = 1;
        while (c + r < nb && ((bm >> (c + r)) & 1u) && r < 8) ++r;
        const uint32_t rank = (uint32_t)__popc(bm & ((1u << c) - 1u));
        runs[nr >> 1] |= (0x80u | ((uint32_t)(r - 1) << 4) | (uint32_t)c | (rank << 8)) << (16 * (nr & 1));
        ++nr;
        c += r;
      }
      P.m_runs[(size_t)cg * nkb + kb] = make_uint4(runs[0], runs[1], runs[2], runs[3]);
    }
  }
}

__device__ __forceinline__ void tmem_zero_32cols(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
    "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr),
    "r"(z)
    : "memory");
}

__global__ void __launch_bounds__(BT_THREADS, 1)
  smm_bf16_planned_kernel(BtPlanPtrs P, float* __restrict__ c_data, const int* __restrict__ c_off, int nrb, int ncb, int nkb, int m, int n, int flags) {
  const int nb = (flags & BT_FLAG_A_TMEM) ? BT_NB_A_TMEM : BT_NB;
  extern __shared__ __align__(1024) unsigned char bt_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const BtGeom g = bt_geom(m, n);
  const int n_rg = (nrb + g.bpt - 1) / g.bpt, n_cg = (ncb + nb - 1) / nb;
  const int n_tiles = bt_num_tiles(n_rg, n_cg);

  uint64_t* full = reinterpret_cast<uint64_t*>(bt_smem);  // [BP_NS]
  uint64_t* empty = full + BP_NS;                         // [BP_NS]
  uint64_t* tmem_full = empty + BP_NS;                    // [1]
  uint64_t* tmem_empty = tmem_full + 1;                   // [1]
  uint32_t* b_start = reinterpret_cast<uint32_t*>(tmem_empty + 1);  // [BP_NS] first ring slot of the stage's packed B blocks
  uint32_t* tmem_ptr = b_start + BP_NS;
  unsigned char* a_stages = bt_smem + 1024;                           // BP_NS x 8 KB
  unsigned char* b_ring = a_stages + (size_t)BP_NS * BT_A_BYTES;      // BP_RING x 2 KB

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the plan kernel (and pack kernels / map uploads) are complete and visible

  // A slots start as zeros (the plan's zero copies rely on it for row group 15 and for slots never written)
  for (size_t i = (size_t)threadIdx.x * 16; i < (size_t)BP_NS * BT_A_BYTES; i += (size_t)BT_THREADS * 16)
    *reinterpret_cast<uint4*>(a_stages + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < BP_NS; ++s) {
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
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;
  if (warp >= 2 && warp < 6) {  // accumulators start as zeros (and are re-zeroed by the epilogue): every MMA accumulates
    const int q = warp & 3;
    for (int c = 0; c < BT_NB; ++c) tmem_zero_32cols(tmem_base + ((uint32_t)(q * 32) << 16) + 32u * (uint32_t)c);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == 6) {
    // ===================================== A producer: one ready-made copy command per lane, fixed 8 KB slot per stage entry ======
    const bool active = lane < BP_RA;
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      int rg, cg;
      bt_tile_coords(t, n_rg, n_cg, rg, cg);
      const uint4* __restrict__ p = P.a_cmd + (size_t)rg * nkb * BP_RA + (active ? lane : 0);
      uint4 cur[BP_KC], nxt[BP_KC];
#pragma unroll
      for (int j = 0; j < BP_KC; ++j) cur[j] = (active && j < nkb) ? __ldg(p + (size_t)j * BP_RA) : make_uint4(0, 0, 0, 0);
      for (int k0 = 0; k0 < nkb; k0 += BP_KC) {
#pragma unroll
        for (int j = 0; j < BP_KC; ++j)
          nxt[j] = (active && k0 + BP_KC + j < nkb) ? __ldg(p + (size_t)(k0 + BP_KC + j) * BP_RA) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < BP_KC; ++j) {
          if (k0 + j < nkb) {  // warp-uniform
            const uint4 cmd = cur[j];
            mbar_wait(&empty[s], ph ^ 1u);
            if (lane == 0) mbar_expect_tx(&full[s], cmd.w);  // this warp's arrival; the phase completes when both warps' bytes have landed
            __syncwarp();
            const uint32_t bytes = cmd.z >> 16;
            if (bytes != 0)
              bulk_g2s(a_stages + (size_t)s * BT_A_BYTES + (cmd.z & 0xffffu),
                       reinterpret_cast<const unsigned char*>((unsigned long long)cmd.x | ((unsigned long long)cmd.y << 32)), bytes, &full[s]);
            if (++s == BP_NS) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < BP_KC; ++j) cur[j] = nxt[j];
      }
    }
  }
  else if (warp == 0) {
    // ===================================== B producer: packed blocks in a ring of 2 KB slots =====================================
    // lane j < BP_NS remembers the ring interval and the iteration of stage entry j; a new stage takes the slots behind the newest
    // one (from slot 0 again once the nominal end is passed) after every older stage that still owns one of them has been consumed
    const bool active = lane < BP_RB;
    int my_start = 0, my_cnt = 0, my_it = -1;
    int it = 0, tail = 0, head = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      int rg, cg;
      bt_tile_coords(t, n_rg, n_cg, rg, cg);
      const uint4* __restrict__ p = P.b_cmd + (size_t)cg * nkb * BP_RB + (active ? lane : 0);
      uint4 cur[BP_KC], nxt[BP_KC];
#pragma unroll
      for (int j = 0; j < BP_KC; ++j) cur[j] = (active && j < nkb) ? __ldg(p + (size_t)j * BP_RB) : make_uint4(0, 0, 0, 0);
      for (int k0 = 0; k0 < nkb; k0 += BP_KC) {
#pragma unroll
        for (int j = 0; j < BP_KC; ++j)
          nxt[j] = (active && k0 + BP_KC + j < nkb) ? __ldg(p + (size_t)(k0 + BP_KC + j) * BP_RB) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < BP_KC; ++j) {
          if (k0 + j < nkb) {  // warp-uniform
            const uint4 cmd = cur[j];
            const uint32_t total = __shfl_sync(0xffffffffu, cmd.w, 0);
            const int cnt = (int)(total / (uint32_t)BT_B_SLOT);
            const int s = it % BP_NS;
            const int start = head >= BP_RING_NOM ? 0 : head;
            const bool ov = lane < BP_NS && my_it >= tail && cnt > 0 && my_cnt > 0 && start < my_start + my_cnt && my_start < start + cnt;
            const int newest = __reduce_max_sync(0xffffffffu, ov ? my_it : -1);
            const int until = max(newest, it - BP_NS);  // plus the stage entry itself
            while (tail <= until) {
              mbar_wait(&empty[tail % BP_NS], (uint32_t)((tail / BP_NS) & 1));
              ++tail;
            }
            if (lane == s) {
              my_start = start;
              my_cnt = cnt;
              my_it = it;
            }
            if (lane == 0) {
              b_start[s] = (uint32_t)start;
              mbar_expect_tx(&full[s], total);
            }
            __syncwarp();
            const uint32_t bytes = cmd.z >> 16;
            if (bytes != 0)
              bulk_g2s(b_ring + (size_t)start * BT_B_SLOT + (cmd.z & 0xffffu),
                       reinterpret_cast<const unsigned char*>((unsigned long long)cmd.x | ((unsigned long long)cmd.y << 32)), bytes, &full[s]);
            head = start + cnt;
            ++it;
          }
        }
#pragma unroll
        for (int j = 0; j < BP_KC; ++j) cur[j] = nxt[j];
      }
    }
  }
  else if (warp == 1) {
    // ===================================== MMA issuer: walks the packed runs of the plan =====================================
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const bool a_tmem = (flags & BT_FLAG_A_TMEM) != 0;
    const uint64_t adesc_base = umma_desc(smem_u32(a_stages), 128u, 512u), bdesc_base = umma_desc(smem_u32(b_ring), 128u, 512u);
    uint32_t it = 0, tile_no = 0, ph = 0;
    int s = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_no) {
      int rg, cg;
      bt_tile_coords(t, n_rg, n_cg, rg, cg);
      const uint4* __restrict__ pr = P.m_runs + (size_t)cg * nkb;
      const unsigned char* __restrict__ pa = P.a_any + (size_t)rg * nkb;
      mbar_wait(tmem_empty, (tile_no & 1u) ^ 1u);  // the epilogue has drained (and zeroed) the previous tile's accumulators
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint4 cur[BP_KC], nxt[BP_KC];
#pragma unroll
      for (int j = 0; j < BP_KC; ++j) {
        const bool in = j < nkb;
        const unsigned char any = in ? __ldg(pa + j) : (unsigned char)0;
        const uint4 rr = in ? __ldg(pr + j) : make_uint4(0, 0, 0, 0);
        cur[j] = any ? rr : make_uint4(0, 0, 0, 0);
      }
      for (int k0 = 0; k0 < nkb; k0 += BP_KC) {
#pragma unroll
        for (int j = 0; j < BP_KC; ++j) {  // both loads independent of each other
          const bool in = k0 + BP_KC + j < nkb;
          const unsigned char any = in ? __ldg(pa + k0 + BP_KC + j) : (unsigned char)0;
          const uint4 rr = in ? __ldg(pr + k0 + BP_KC + j) : make_uint4(0, 0, 0, 0);
          nxt[j] = any ? rr : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < BP_KC; ++j) {
          if (k0 + j < nkb) {
            mbar_wait(&full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
              unsigned long long lo = (unsigned long long)cur[j].x | ((unsigned long long)cur[j].y << 32);
              unsigned long long hi = (unsigned long long)cur[j].z | ((unsigned long long)cur[j].w << 32);
              const uint64_t ad0 = adesc_base + (uint64_t)((uint32_t)s * (uint32_t)(BT_A_BYTES >> 4));
              const uint64_t bd0 = bdesc_base + (uint64_t)(b_start[s] * (uint32_t)(BT_B_SLOT >> 4));
              const uint32_t ta = tmem_base + (uint32_t)(BT_NB_A_TMEM * 32) + 16u * (it & 1u);
              if (a_tmem && (lo & 0x80ull)) {
                utccp_128x256b(ta, ad0);
                utccp_128x256b(ta + 8u, ad0 + 16u);
              }
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                unsigned long long runs = half == 0 ? lo : hi;
                while (runs & 0x80ull) {
                  const uint32_t w = (uint32_t)runs;
                  const uint32_t c0 = w & 15u, r = ((w >> 4) & 7u) + 1u, rank = (w >> 8) & 15u;
                  const uint32_t idesc = idesc_base | ((4u * r) << 17);                     // N = 32 r
                  const uint64_t bd = bd0 + (uint64_t)(rank * (uint32_t)(BT_B_SLOT >> 4));  // start address in 16-byte units
                  const uint32_t d = tmem_base + 32u * c0;
                  if (a_tmem) {
                    umma_bf16_ts(d, ta, bd, idesc, 1u);
                    umma_bf16_ts(d, ta + 8u, bd + 16u, idesc, 1u);
                  }
                  else {
                    umma_bf16(d, ad0, bd, idesc, 1u);
                    umma_bf16(d, ad0 + 16u, bd + 16u, idesc, 1u);  // k = 16..31: +256 B
                  }
                  runs >>= 16;
                }
              }
              umma_commit(&empty[s]);  // the stage may be refilled once these MMAs have read it
            }
            __syncwarp();
            ++it;
            if (++s == BP_NS) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < BP_KC; ++j) cur[j] = nxt[j];
      }
      if (elect_one()) umma_commit(tmem_full);
      __syncwarp();
    }
  }
  else if (warp < 6) {
    // ===================================== epilogue (4 warps = 128 TMEM lanes) =====================================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rg_rows = g.rg_a * 8;
    const int blk = row / rg_rows, r_in = row - blk * rg_rows;
    uint32_t tile_no = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_no) {
      int rg, cg;
      bt_tile_coords(t, n_rg, n_cg, rg, cg);
      const int rb = rg * g.bpt + blk;
      const bool row_ok = blk < g.bpt && r_in < m && rb < nrb;
      mbar_wait(tmem_full, tile_no & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = 0; c < nb; ++c) {
        const int cb = cg * nb + c;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 32u * (uint32_t)c;
        uint32_t r[32];
        asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
            "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
            "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
            "=r"(r[31])
          : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tmem_zero_32cols(taddr);  // ready for the next tile
        if (row_ok && cb < ncb) {
          const int off = __ldg(c_off + (size_t)rb * ncb + cb);
          if (off >= 0) {
            float* __restrict__ dst = c_data + (size_t)off + r_in;
#pragma unroll
            for (int col = 0; col < 32; ++col)
              if (col < n) dst[(size_t)col * m] = __uint_as_float(r[col]);
          }
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
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
