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

// Third step (measured, B200): with the producers' bookkeeping gone the kernel ran at 745 cycles per k block, and packing the B blocks
// into a ring with 8 stage entries (1.6x the bytes in flight) made it SLOWER (43 vs 34 ms) -- the limit is not the copy latency but
// the MMA issuer: SASS showed ~24 uniform-datapath instructions per run of blocks (unpacking the run, three 64-bit descriptor
// additions, two predicated MMA forms) in a loop with a carried dependency, 6-7 cycles each.  So the plan now carries, per run, the
// finished 32-bit words {B descriptor offset | accumulator column offset, instruction descriptor}; the B producer's spare lanes
// copy the 64 bytes of a k block's runs into shared memory beside the stage, and the issuer reads them with four LDS.128 and
// issues an unrolled, nested sequence (run q+1 is only looked at when run q exists): no global loads, no loop-carried chain.
// Fourth step: the pipeline depth is not the limit either (5 / 4 / 3 stages: 28.8 / 29.5 / 33.0 ms); the issuer still executes ~90
// instructions per k block.  Its work is now split over TWO issuer warps that own disjoint halves of the accumulator (block columns
// 0-7 | 8-15; a run never crosses the middle): both read the same operand stage, each issues and commits its own MMAs (independent
// TMEM columns, so no ordering between the two threads is needed), a stage is free when both have committed.
// (Measured and dropped, in the history of this file: A operand copied to TMEM once per k block with FOUR issuer warps, 211 TFLOP/s --
// slower than the two-issuer shared-memory version below, 234.)
constexpr int BP_NS = BT_STAGES;  // stage entries: fixed slots, as in smm_bf16_tiled.cuh
constexpr int BP_THREADS = 256;   // warp 0: B producer, 1 and 7: MMA issuers, 2-5: epilogue, 6: A producer
constexpr int BP_RA = 8;   // command slots per row of the A plan (<= 5 used)
constexpr int BP_RB = 20;  // uint4 per row of the B plan: 16 copy commands + 4 x (2 MMA runs)
constexpr int BP_KC = 8;   // plan rows prefetched per lane

struct BtPlanPtrs {
  uint4* a_cmd;           // [n_rg][nkb][BP_RA]: x,y = source address, z = stage offset | bytes << 16, w (slot 0) = bytes of the row
  uint4* b_cmd;           // [n_cg][nkb][BP_RB]: 16 copy commands, then 8 MMA runs as {B descriptor offset (16-byte units) | TMEM column
                          // offset << 16, instruction descriptor (0 = no run)}
  unsigned char* a_any;   // [n_rg][nkb]: the row group has an A block in this k block
  unsigned char* zeros;   // 5 A tiles of zeros
};
inline size_t bp_smem_bytes(int ns) { return 1024 + (size_t)ns * (BT_A_BYTES + BT_NB * BT_B_SLOT); }
inline size_t bt_plan_bytes(int n_rg, int n_cg, int nkb, size_t* off /* [5] */) {
  size_t o = 0;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  off[0] = o, o += up((size_t)n_rg * nkb * BP_RA * sizeof(uint4));
  off[1] = o, o += up((size_t)n_cg * nkb * BP_RB * sizeof(uint4));
  off[2] = o;
  off[3] = o, o += up((size_t)n_rg * nkb);
  off[4] = o, o += 5 * 2048;
  return o;
}

__device__ __forceinline__ uint4 bt_cmd(const unsigned char* src, uint32_t dst, uint32_t bytes) {
  const unsigned long long a = (unsigned long long)src;
  return make_uint4((uint32_t)a, (uint32_t)(a >> 32), dst | (bytes << 16), 0u);
}

__global__ void bt_plan_kernel(const unsigned char* __restrict__ a_tiles, const int* __restrict__ a_map, const unsigned char* __restrict__ b_tiles,
                               const int* __restrict__ b_map, int nrb, int ncb, int nkb, int m, int n, int nb, int ns, BtPlanPtrs P) {
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
        int kp = kb - ns;
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
            kp -= ns;
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
      int c = 0;
      while (c < nb) {  // copy commands: adjacent existing blocks whose tiles are adjacent in memory, into their fixed slots
        if (first_idx[c] < 0) {
          ++c;
          continue;
        }
        int len = 1;
        while (c + len < nb && first_idx[c + len] == first_idx[c] + len) ++len;
        const uint32_t bytes = (uint32_t)(len * BT_B_SLOT);
        out[nc++] = bt_cmd(b_tiles + (size_t)first_idx[c] * BT_B_SLOT, (uint32_t)(BT_A_BYTES + c * BT_B_SLOT), bytes);
        total_b += bytes;
        c += len;
      }
      for (int q = nc; q < 16; ++q) out[q] = make_uint4(0, 0, 0, 0);
      out[0].w = total_b;
      // MMA runs: adjacent existing blocks, N = 32 * run <= 256; cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format
      // BF16 (1) [7,10),[10,13), K-major A and B, N >> 3 at [17,23), M >> 4 at [24,29)
      uint32_t rec[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) rec[q] = 0u;
      for (int half = 0; half < 2; ++half) {  // words 0..7: block columns 0-7 (first issuer), 8..15: block columns 8-15 (second issuer)
        int nr = 0;
        c = 8 * half;
        const int cend = min(nb, 8 * half + 8);
        while (c < cend) {
          if (!((bm >> c) & 1u)) {
            ++c;
            continue;
          }
          int r = 1;
          while (c + r < cend && ((bm >> (c + r)) & 1u)) ++r;
          rec[8 * half + 2 * nr] = (uint32_t)(c * (BT_B_SLOT >> 4)) | ((uint32_t)(32 * c) << 16);
          rec[8 * half + 2 * nr + 1] = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24) | ((uint32_t)(4 * r) << 17);
          ++nr;
          c += r;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) out[16 + q] = make_uint4(rec[4 * q], rec[4 * q + 1], rec[4 * q + 2], rec[4 * q + 3]);
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

// one run of adjacent existing B blocks: two MMAs (k = 0..15 | 16..31) into the accumulators of its block columns
#define BP_RUN(X, Y)                                                                                                       \
  {                                                                                                                        \
    const uint32_t bl_ = bd_lo + ((X) & 0xffffu), d_ = tmem_base + ((X) >> 16);                                            \
    umma_bf16(d_, ((uint64_t)ad_hi << 32) | ad_lo, ((uint64_t)bd_hi << 32) | bl_, (Y), 1u);                                \
    umma_bf16(d_, ((uint64_t)ad_hi << 32) | (ad_lo + 16u), ((uint64_t)bd_hi << 32) | (bl_ + 16u), (Y), 1u);                \
  }

__global__ void __launch_bounds__(BP_THREADS, 1)
  smm_bf16_planned_kernel(BtPlanPtrs P, float* __restrict__ c_data, const int* __restrict__ c_off, int nrb, int ncb, int nkb, int m, int n, int ns) {
  const int nb = BT_NB;
  extern __shared__ __align__(1024) unsigned char bt_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const BtGeom g = bt_geom(m, n);
  const int n_rg = (nrb + g.bpt - 1) / g.bpt, n_cg = (ncb + nb - 1) / nb;
  const int n_tiles = bt_num_tiles(n_rg, n_cg);

  uint64_t* full = reinterpret_cast<uint64_t*>(bt_smem);  // [BP_NS]
  uint64_t* empty = full + BP_NS;                         // [BP_NS]
  uint64_t* tmem_full = empty + BP_NS;                    // [1]
  uint64_t* tmem_empty = tmem_full + 1;                   // [1]
  uint32_t* a_flag = reinterpret_cast<uint32_t*>(tmem_empty + 1);  // [BP_NS] the stage has an A block
  uint32_t* tmem_ptr = a_flag + BP_NS;
  uint4* run_words = reinterpret_cast<uint4*>(bt_smem + 512);      // [BP_NS][4]: the stage's MMA runs (8 x {offsets, idesc})
  unsigned char* stages = bt_smem + 1024;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the plan kernel (and pack kernels / map uploads) are complete and visible

  for (size_t i = (size_t)threadIdx.x * 16; i < (size_t)ns * g.stage; i += (size_t)BP_THREADS * 16)
    *reinterpret_cast<uint4*>(stages + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < BP_NS; ++s) {
      mbar_init(&full[s], 2);   // the two producer warps
      mbar_init(&empty[s], 2);  // the two issuer warps
    }
    mbar_init(tmem_full, 2);
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

  if (warp == 0 || warp == 6) {
    // ===================================== TMA producers: one ready-made copy command per lane =====================================
    // warp 6: A blocks (+ zero copies where a slot may hold stale data); warp 0: B blocks, and its lanes 16..19 pass the k block's
    // MMA run words on to the issuer through shared memory
    const bool a_role = warp == 6;
    const int R = a_role ? BP_RA : BP_RB;
    const bool active = lane < R;
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      int rg, cg;
      bt_tile_coords(t, n_rg, n_cg, rg, cg);
      const uint4* __restrict__ p = (a_role ? P.a_cmd + (size_t)rg * nkb * BP_RA : P.b_cmd + (size_t)cg * nkb * BP_RB) + (active ? lane : 0);
      uint4 cur[BP_KC], nxt[BP_KC];
#pragma unroll
      for (int j = 0; j < BP_KC; ++j) cur[j] = (active && j < nkb) ? __ldg(p + (size_t)j * R) : make_uint4(0, 0, 0, 0);
      for (int k0 = 0; k0 < nkb; k0 += BP_KC) {
#pragma unroll
        for (int j = 0; j < BP_KC; ++j) nxt[j] = (active && k0 + BP_KC + j < nkb) ? __ldg(p + (size_t)(k0 + BP_KC + j) * R) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < BP_KC; ++j) {
          if (k0 + j < nkb) {  // warp-uniform
            const uint4 cmd = cur[j];
            mbar_wait(&empty[s], ph ^ 1u);
            if (!a_role && lane >= 16 && lane < 20) run_words[s * 4 + (lane - 16)] = cmd;
            __syncwarp();
            if (lane == 0) {
              if (a_role) a_flag[s] = cmd.w;
              mbar_expect_tx(&full[s], cmd.w);  // this warp's arrival (release); the phase completes when both warps' bytes have landed
            }
            const uint32_t bytes = cmd.z >> 16;
            if (lane < 16 && bytes != 0)
              bulk_g2s(stages + (size_t)s * g.stage + (cmd.z & 0xffffu),
                       reinterpret_cast<const unsigned char*>((unsigned long long)cmd.x | ((unsigned long long)cmd.y << 32)), bytes, &full[s]);
            if (++s == ns) {
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
  else if (warp == 1 || warp == 7) {
    // ===================================== MMA issuers: warp 1 block columns 0-7, warp 7 block columns 8-15 =====================
    const int half = warp == 1 ? 0 : 1;
    const uint64_t adesc_base = umma_desc(smem_u32(stages), 128u, 512u), bdesc_base = umma_desc(smem_u32(stages) + (uint32_t)BT_A_BYTES, 128u, 512u);
    const uint32_t ad_hi = (uint32_t)(adesc_base >> 32), bd_hi = (uint32_t)(bdesc_base >> 32);
    const uint32_t stage16 = (uint32_t)g.stage >> 4;
    uint32_t tile_no = 0, ph = 0;
    int s = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_no) {
      mbar_wait(tmem_empty, (tile_no & 1u) ^ 1u);  // the epilogue has drained (and zeroed) the previous tile's accumulators
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint4 q0 = run_words[s * 4 + 2 * half];
          if (a_flag[s] != 0u && q0.y != 0u) {
            const uint4 q1 = run_words[s * 4 + 2 * half + 1];
            const uint32_t ad_lo = (uint32_t)adesc_base + (uint32_t)s * stage16, bd_lo = (uint32_t)bdesc_base + (uint32_t)s * stage16;
            BP_RUN(q0.x, q0.y)
            if (q0.w != 0u) {
              BP_RUN(q0.z, q0.w)
              if (q1.y != 0u) {
                BP_RUN(q1.x, q1.y)
                if (q1.w != 0u) BP_RUN(q1.z, q1.w)
              }
            }
          }
          umma_commit(&empty[s]);  // the stage may be refilled once the MMAs of BOTH issuers have read it
        }
        __syncwarp();
        if (++s == ns) {
          s = 0;
          ph ^= 1u;
        }
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace smm
