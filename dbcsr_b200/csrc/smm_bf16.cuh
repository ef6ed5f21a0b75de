// dbcsr_b200/csrc/smm_bf16.cuh -- BF16 stack drain on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// EXTENSION of the DBCSR ABI (dbcsr_type_bf16_ext = 9; DBCSR itself has only real_4/8 and complex_4/8,
// src/acc/acc_libsmm.h:31-36): A and B panels are converted once per panel from FP64 to BF16 *tiles* in the canonical UMMA
// operand layout (pack_bf16_kernel), C is accumulated in FP32.  BASELINE.json config 4 (23x23 blocks, 50 % occupation).
//
// Tile format (one per block, 128-byte aligned, element (row, kk) of a ROWS x KDIM operand):
//     byte offset = (kk / 8) * (RG * 128) + (row / 8) * 128 + (row % 8) * 16 + (kk % 8) * 2,   RG = ceil(ROWS / 8)
// i.e. K-major "interleave / no swizzle" core matrices of 8 rows x 8 k (128 B), row groups contiguous inside a k group.
// That is exactly what a tcgen05 shared-memory descriptor with SBO = 128 B and LBO = RG*128 B describes, so a tile is staged
// with ONE cp.async.bulk and fed to the MMA without any shuffling; padding rows/k are stored as zeros by the pack kernel.
//
// Kernel structure (one CTA = 4 warps, several CTAs per SM share TMEM 128 columns each):
//   warp 1        : TMA producers - lane s owns ring slot s: per stack entry two bulk copies (A tile, B tile)
//   warp 2 lane 0 : MMA issuer    - per entry ceil(K/16) tcgen05.mma (M=128, N=32, K=16, D in TMEM); a run of equal c_first
//                                    accumulates in one TMEM accumulator; tcgen05.commit releases the slot / publishes the run
//   warp 0        : epilogue      - tcgen05.ld of the finished accumulator (row = lane), RED.ADD.F32 into the C block
// Rows >= m of the 128-row MMA operand and columns >= n read neighbouring shared memory: they only produce accumulator
// rows/columns that are never stored.  k padding is zero in both operands, so it contributes nothing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "smm_dmma.cuh"

namespace smm {

constexpr int BF_SLOTS = 8;    // ring slots (entries in flight) per CTA
constexpr int BF_ACC = 2;      // TMEM accumulators per CTA, 32 columns each (64 columns => up to 8 CTAs per SM)
constexpr int BF_TMEM_COLS = BF_ACC * 32;
constexpr int BF_THREADS = 128;

struct Bf16Geom {
  int rg_a, rg_b, kg, kg_slot;       // row groups of A / B tiles, k groups per tile, k groups per slot (even)
  int tile_a, tile_b, slot_a, slot_b;  // bytes
};

__host__ __device__ inline Bf16Geom bf16_geom(int m, int n, int k) {
  Bf16Geom g;
  g.rg_a = (m + 7) / 8;
  g.rg_b = (n + 7) / 8;
  g.kg = (k + 7) / 8;
  g.kg_slot = (g.kg + 1) & ~1;
  g.tile_a = g.kg * g.rg_a * 128;
  g.tile_b = g.kg * g.rg_b * 128;
  g.slot_a = g.kg_slot * g.rg_a * 128;
  g.slot_b = g.kg_slot * g.rg_b * 128;
  return g;
}

__host__ __device__ inline size_t bf16_smem_bytes(const Bf16Geom& g) {
  // barriers + bookkeeping (512 B) | A slots | B slots | slack for the 128-row / 32-column over-read of the last slot
  return 512 + (size_t)BF_SLOTS * (g.slot_a + g.slot_b) + (size_t)g.kg_slot * 2048 + 2048;
}

// FP64 block (element (row,kk) at src[row*row_stride + kk*k_stride]) -> BF16 tile (round to nearest even), one warp per block.
__global__ void pack_bf16_kernel(const double* __restrict__ src, int nblks, int rows, int kdim, int row_stride, int k_stride,
                                 unsigned char* __restrict__ dst) {
  const int wpc = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = (rows + 7) / 8, kg = (kdim + 7) / 8;
  const int tile_bytes = kg * rg * 128;
  const int nelem = kg * rg * 64;  // bf16 elements per tile incl. padding
  for (int b = blockIdx.x * wpc + warp; b < nblks; b += gridDim.x * wpc) {
    const double* __restrict__ s = src + (size_t)b * rows * kdim;
    unsigned short* __restrict__ d = reinterpret_cast<unsigned short*>(dst + (size_t)b * tile_bytes);
    for (int i = lane; i < nelem; i += 32) {
      const int kk8 = i & 7, r8 = (i >> 3) & 7, rgi = (i >> 6) % rg, kgi = (i >> 6) / rg;
      const int row = rgi * 8 + r8, kk = kgi * 8 + kk8;
      float v = 0.f;
      if (row < rows && kk < kdim) v = (float)s[(size_t)row * row_stride + (size_t)kk * k_stride];
      // round-to-nearest-even bf16 (values are finite)
      unsigned int u = __float_as_uint(v);
      u += 0x7fffu + ((u >> 16) & 1u);
      d[i] = (unsigned short)(u >> 16);
    }
  }
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start address [0,14) >> 4, LBO [16,30) >> 4, SBO [32,46) >> 4,
  // version = 1 at [46,48), layout type SWIZZLE_NONE = 0 at [61,64)
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// mbar_arrive: smm_dmma.cuh
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "setp.ne.b32 p, %4, 0;\n"
    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
    "}\n" ::"r"(tmem_d),
    "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
    : "memory");
}

__global__ void __launch_bounds__(BF_THREADS) smm_bf16_kernel(const int* __restrict__ stack, int stack_size,
                                                               const unsigned char* __restrict__ a_tiles,
                                                               const unsigned char* __restrict__ b_tiles, float* __restrict__ c_data, int m,
                                                               int n, int k, int chunk) {
  extern __shared__ __align__(128) unsigned char bf_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e0 = blockIdx.x * chunk;
  const int e1 = min(e0 + chunk, stack_size);
  if (e0 >= e1) {  // whole CTA
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    return;
  }
  const int nent = e1 - e0;
  const Bf16Geom g = bf16_geom(m, n, k);
  const int mk = m * k, nk = n * k;

  uint64_t* full = reinterpret_cast<uint64_t*>(bf_smem);          // [BF_SLOTS]
  uint64_t* empty = full + BF_SLOTS;                                // [BF_SLOTS]
  uint64_t* acc_full = empty + BF_SLOTS;                            // [BF_ACC]
  uint64_t* acc_empty = acc_full + BF_ACC;                          // [BF_ACC]
  int* acc_c = reinterpret_cast<int*>(acc_empty + BF_ACC);          // [BF_ACC] c_first of the run held by the accumulator
  int* slot_c = acc_c + BF_ACC;                                     // [BF_SLOTS] c_first of the entry in the slot
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(slot_c + BF_SLOTS);
  unsigned char* slots_a = bf_smem + 512;
  unsigned char* slots_b = slots_a + (size_t)BF_SLOTS * g.slot_a;
  const size_t ring_bytes = (size_t)BF_SLOTS * (g.slot_a + g.slot_b) + (size_t)g.kg_slot * 2048 + 2048;

  // Programmatic dependent launch: become resident early, but wait for the predecessor (pack_bf16, memset, another drain) to be
  // complete and visible before anything is read -- and before TMEM is allocated, so that a waiting dependent CTA can never hold
  // TMEM columns a predecessor CTA still needs.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // The k padding group of every slot must read as zero (tiles carry kg groups, a slot kg_slot): zero just those bytes and the
  // slack behind the ring once; everything else is either overwritten by the TMA or only feeds discarded rows/columns.
  {
    const int pad_a = g.slot_a - g.tile_a, pad_b = g.slot_b - g.tile_b;
    for (int s = 0; s < BF_SLOTS; ++s) {
      for (int i = threadIdx.x * 16; i < pad_a; i += BF_THREADS * 16)
        *reinterpret_cast<uint4*>(slots_a + (size_t)s * g.slot_a + g.tile_a + i) = make_uint4(0, 0, 0, 0);
      for (int i = threadIdx.x * 16; i < pad_b; i += BF_THREADS * 16)
        *reinterpret_cast<uint4*>(slots_b + (size_t)s * g.slot_b + g.tile_b + i) = make_uint4(0, 0, 0, 0);
    }
    const size_t ring_only = (size_t)BF_SLOTS * (g.slot_a + g.slot_b);
    for (size_t i = ring_only + threadIdx.x * 16; i < ring_bytes; i += BF_THREADS * 16)
      *reinterpret_cast<uint4*>(slots_a + i) = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < BF_SLOTS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < BF_ACC; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(BF_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic zero-fill -> visible to TMA / tensor-core proxies
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 1) {
    // ===== TMA producers: lane s owns ring slot s and the entries i = s, s + SLOTS, ... (issue latencies overlap) =====
    if (lane < BF_SLOTS) {
      const int s = lane;
      for (int i = lane; i < nent; i += BF_SLOTS) {
        const int3 p = ld_entry(stack, e0 + i);
        mbar_wait(&empty[s], (uint32_t)(((i / BF_SLOTS) & 1) ^ 1));
        slot_c[s] = p.z;
        mbar_expect_tx(&full[s], (uint32_t)(g.tile_a + g.tile_b));
        bulk_g2s(slots_a + (size_t)s * g.slot_a, a_tiles + (size_t)((p.x - 1) / mk) * g.tile_a, (uint32_t)g.tile_a, &full[s]);
        bulk_g2s(slots_b + (size_t)s * g.slot_b, b_tiles + (size_t)((p.y - 1) / nk) * g.tile_b, (uint32_t)g.tile_b, &full[s]);
      }
    }
  }
  else if (warp == 2) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format BF16 (1) [7,10),[10,13), K-major A and B,
      // N >> 3 at [17,23), M >> 4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t lbo_a = (uint32_t)g.rg_a * 128u, lbo_b = (uint32_t)g.rg_b * 128u;
      int cur_c = -1, run = -1, acc = 0;
      uint32_t accumulate = 0;
      for (int i = 0; i < nent; ++i) {
        const int s = i % BF_SLOTS;
        mbar_wait(&full[s], (uint32_t)((i / BF_SLOTS) & 1));
        const int c = slot_c[s];
        if (c != cur_c) {
          if (run >= 0) umma_commit(&acc_full[acc]);  // everything issued so far has to finish before the epilogue reads
          ++run;
          acc = run % BF_ACC;
          mbar_wait(&acc_empty[acc], (uint32_t)(((run / BF_ACC) & 1) ^ 1));
          acc_c[acc] = c;
          __threadfence_block();  // the epilogue reads acc_c after the (asynchronous) commit-arrive on acc_full
          cur_c = c;
          accumulate = 0;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(slots_a + (size_t)s * g.slot_a), sb = smem_u32(slots_b + (size_t)s * g.slot_b);
        for (int kk = 0; kk < g.kg_slot / 2; ++kk) {
          umma_bf16(tmem_base + (uint32_t)acc * 32u, umma_desc(sa + (uint32_t)kk * 2u * lbo_a, lbo_a, 128u),
                    umma_desc(sb + (uint32_t)kk * 2u * lbo_b, lbo_b, 128u), idesc, accumulate);
          accumulate = 1;
        }
        umma_commit(&empty[s]);  // slot may be refilled once these MMAs have read it
      }
      umma_commit(&acc_full[acc]);
      // sentinel run: tells the epilogue warp to stop
      ++run;
      acc = run % BF_ACC;
      mbar_wait(&acc_empty[acc], (uint32_t)(((run / BF_ACC) & 1) ^ 1));
      acc_c[acc] = -1;
      __threadfence_block();
      mbar_arrive(&acc_full[acc]);
    }
  }
  else if (warp == 0) {
    // ===== epilogue: TMEM lanes 0..31 belong to warp 0 =====
    for (int run = 0;; ++run) {
      const int acc = run % BF_ACC;
      mbar_wait(&acc_full[acc], (uint32_t)((run / BF_ACC) & 1));
      const int c = acc_c[acc];
      if (c < 0) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32];
      asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(tmem_base + (uint32_t)acc * 32u));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (lane < m) {
        float* cb = c_data + (c - 1) + lane;
#pragma unroll
        for (int col = 0; col < 32; ++col)
          if (col < n) atomicAdd(cb + (size_t)col * m, __uint_as_float(r[col]));
      }
    }
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BF_TMEM_COLS) : "memory");
  }
}

}  // namespace smm
