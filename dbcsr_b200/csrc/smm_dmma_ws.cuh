// dbcsr_b200/csrc/smm_dmma_ws.cuh -- warp-specialised FP64 stack-drain kernel for sm_100a (producer/consumer variant of
// smm_dmma.cuh; same arithmetic, same operand staging, same RED flush, different division of labour).
//
// Why: in the warp-autonomous kernel every warp alternates between issuing its own TMA copies (single-lane address code), waiting
// for them and feeding the FP64 tensor pipe; the shared-memory budget (about 24 entry stages per SM for 23x23 blocks) is what
// bounds the number of entries in flight, whatever the (warps, stages) split (profiles/variants_r01_*.txt).  Here the roles are
// separated inside one CTA:
//   * NC consumer warps do nothing but wait on a `full` mbarrier, load DMMA fragments from the stage and issue DMMA.8x8x4;
//     consumer c owns a contiguous chunk of the C-sorted stack (so the run accumulation in registers is kept) and a private
//     ring of D stages, i.e. its operands are prefetched D-1 entries ahead;
//   * ONE producer warp serves all NC*D stages, one stage per LANE: lane (c,d) waits (non-blocking test) for `empty[c][d]`,
//     then issues the two cp.async.bulk copies of entry d, d+D, ... of consumer c.  All lanes run the same loop, so the
//     address arithmetic of several entries is done SIMT-parallel and the producer batches automatically when it falls behind;
//   * the CTA's slice of the parameter stack is copied to shared memory once in the prologue (the only CTA-wide barrier), so
//     neither role has a global load on its per-entry path.
// Split of the stack: consumers are numbered gw = blockIdx.x * NC + c and take floor(S/W) or floor(S/W)+1 entries (W = all
// consumers), cf. warp_chunk() with extra >= 0.
#pragma once
#include "smm_dmma.cuh"

namespace smm {

constexpr int WS_ENT_CAP = 512;  // stack entries per CTA held in shared memory (6 KB); the launcher sizes the grid accordingly

template <int M, int N, int K, int NC, int D>
struct WsGeom {
  using SH = Shape<M, N, K>;
  static constexpr int NSTG = NC * D;
  static constexpr int BAR_BYTES = round_up_c(2 * NSTG * 8, 128);
  static constexpr int ENT_BYTES = round_up_c(3 * WS_ENT_CAP * 4, 128);
  static constexpr int SMEM = BAR_BYTES + ENT_BYTES + NSTG * SH::STAGE;
  static constexpr int THREADS = (NC + 1) * 32;
};

// (consumers, stages per consumer) for a shape: as many stages as one CTA's shared memory holds, at most one producer lane each
template <int M, int N, int K>
struct WsPick {
  static constexpr int FIT = (227 * 1024 - round_up_c(3 * WS_ENT_CAP * 4, 128) - 512) / Shape<M, N, K>::STAGE;
  static constexpr int NC = FIT >= 16 ? 8 : (FIT >= 8 ? 4 : (FIT >= 4 ? 2 : 1));
  static constexpr int D = min_c(3, max_c(1, FIT / NC));
};

// FLUSH: 0 = per-element RED when a run ends (flush_acc); 2 = the finished run is written into the stage the consumer has just
// drained and added to C by one cp.reduce.async.bulk (flush_acc_bulk); that stage goes back to the producer one entry later,
// when the bulk engine has certainly read it.  STAG: the second consumer of every SM sub-partition (warps 4..7 of 8) idles for
// about one entry's DMMA time before its first entry, so that the two consumers of a sub-partition do not run in lock-step
// (both in their inter-entry gap at the same time leaves the FP64 pipe idle; profiles/r01_trace_analysis.md).
template <int M, int N, int K, int NC, int D, int HINT = 0, bool TRACE = false, int FLUSH = 0, bool STAG = false>
__global__ void __launch_bounds__((NC + 1) * 32) smm_dmma_ws_kernel(const int* __restrict__ stack, int stack_size,
                                                                    const double* __restrict__ a_data, const double* __restrict__ b_data,
                                                                    double* __restrict__ c_data, unsigned long long a_limit,
                                                                    unsigned long long b_limit, int base, int extra,
                                                                    unsigned long long* __restrict__ trace) {
  using SH = Shape<M, N, K>;
  using G = WsGeom<M, N, K, NC, D>;
  constexpr int TM = SH::TM, TN = SH::TN, NSTG = G::NSTG;
  static_assert(NSTG <= 32, "one producer lane per stage");
  static_assert(FLUSH != 2 || SH::STAGE >= scratch_bytes(M, N), "the C image of a run must fit into one stage");
  static_assert(FLUSH != 2 || D >= 2, "a stage that is being flushed is released one entry later: needs a second stage");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + NSTG;
  int* ent = reinterpret_cast<int*>(smem_raw + G::BAR_BYTES);
  unsigned char* stages = smem_raw + G::BAR_BYTES + G::ENT_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw0 = blockIdx.x * NC;
  // consumer gw owns [cb(gw), cb(gw+1)); the CTA owns the contiguous range of its NC consumers
  auto cb = [&](int gw) { return min(gw * base + min(gw, extra), stack_size); };
  const int cta_e0 = cb(gw0), cta_e1 = cb(gw0 + NC);
  const int cta_len = cta_e1 - cta_e0;  // <= WS_ENT_CAP by construction of the grid

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the predecessor may have produced A, B, C or the stack: complete + visible first
  if (cta_len <= 0) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    return;
  }

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 3 * cta_len; i += G::THREADS) ent[i] = __ldg(stack + 3 * (size_t)cta_e0 + i);
  __syncthreads();

  if (warp == NC) {
    // ------------------------------------------------ producer: lane = stage ------------------------------------------
    const bool mine = lane < NSTG;
    const int c = lane / D, d = lane - c * D;
    const int e0c = cb(gw0 + c);
    const int lenc = mine ? cb(gw0 + c + 1) - e0c : 0;
    uint64_t pol_ab = 0;
    if (HINT >= 2) pol_ab = policy_evict_last();
    unsigned char* stg = stages + (size_t)(mine ? lane : 0) * SH::STAGE;
    uint64_t* fbar = &full[mine ? lane : 0];
    uint64_t* ebar = &empty[mine ? lane : 0];
    int i = d;
    uint32_t ph = 1;  // a fresh `empty` barrier passes the wait on the preceding phase: all stages start free
    bool active = mine && i < lenc;
    while (__any_sync(0xffffffffu, active)) {
      const bool go = active && mbar_test(ebar, ph);
      if (go) {
        const int* pe = ent + 3 * (e0c - cta_e0 + i);
        const uint64_t ga = reinterpret_cast<uint64_t>(a_data + (pe[0] - 1));
        const uint64_t gb = reinterpret_cast<uint64_t>(b_data + (pe[1] - 1));
        // windows (see stage_block): aligned source, byte counts, and the by-hand tail for the last block of an allocation
        const uint64_t sa = ga & ~15ull, sb = gb & ~15ull;
        const uint32_t sha = (uint32_t)(ga & 15ull), shb = (uint32_t)(gb & 15ull);
        uint32_t ba = (sha + SH::A_BYTES + 15u) & ~15u, bb = (shb + SH::B_BYTES + 15u) & ~15u;
        if (a_limit != 0 && sa + ba > a_limit) {
          ba = (uint32_t)(a_limit - sa) & ~15u;
          for (uint32_t o = ba; o < sha + SH::A_BYTES; o += 8)
            *reinterpret_cast<double*>(stg + o) = *reinterpret_cast<const double*>(sa + o);
        }
        if (b_limit != 0 && sb + bb > b_limit) {
          bb = (uint32_t)(b_limit - sb) & ~15u;
          for (uint32_t o = bb; o < shb + SH::B_BYTES; o += 8)
            *reinterpret_cast<double*>(stg + SH::ABUF + o) = *reinterpret_cast<const double*>(sb + o);
        }
        // arrive (release: the by-hand stores above become visible to the consumer's acquire) with the byte count, then copy
        mbar_expect_tx(fbar, ba + bb);
        if (ba > 0) {
          if (HINT >= 2)
            bulk_g2s_hint(stg, reinterpret_cast<const void*>(sa), ba, fbar, pol_ab);
          else
            bulk_g2s(stg, reinterpret_cast<const void*>(sa), ba, fbar);
        }
        if (bb > 0) {
          if (HINT >= 2)
            bulk_g2s_hint(stg + SH::ABUF, reinterpret_cast<const void*>(sb), bb, fbar, pol_ab);
          else
            bulk_g2s(stg + SH::ABUF, reinterpret_cast<const void*>(sb), bb, fbar);
        }
        i += D;
        ph ^= 1u;
        active = i < lenc;
      }
      if (!__any_sync(0xffffffffu, go)) __nanosleep(32);
    }
  }
  else {
    // ------------------------------------------------ consumers ---------------------------------------------------------
    const int g = lane >> 2, t = lane & 3;
    const int gw = gw0 + warp;
    const int e0 = cb(gw);
    const int len = cb(gw + 1) - e0;
    unsigned long long* rec = nullptr;
    unsigned long long t_flush = 0, n_flush = 0;
    if (TRACE) {
      if (trace != nullptr && gw < 4096 && lane == 0 && len > 0) {
        rec = trace + (size_t)gw * TRACE_WORDS;
        rec[0] = (unsigned long long)smid_now() | ((unsigned long long)len << 32);
        rec[1] = globaltimer_now();
        rec[2] = clock_now();
        rec[3] = (unsigned long long)e0;
      }
    }
    uint64_t pol_c = 0;
    if (HINT >= 1) pol_c = policy_evict_first();

    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int* pe = ent + 3 * (e0 - cta_e0);
    int cur_c = -1;
    int slot = 0;
    uint32_t ph = 0;
    int pending = -1;  // FLUSH == 2: stage still being read by a bulk reduction
    for (int i = 0; i < len; ++i, pe += 3) {
      const int pa = pe[0], pb = pe[1], pc = pe[2];  // shared-memory broadcast loads
      if (FLUSH == 0 && pc != cur_c) {
        if (cur_c >= 0) {
          unsigned long long tf0 = 0;
          if (TRACE) tf0 = clock_now();
          flush_acc<M, N, K, HINT>(c_data, cur_c, acc, g, t, pol_c);
          if (TRACE) {
            t_flush += clock_now() - tf0;
            ++n_flush;
          }
        }
        cur_c = pc;
      }
      const int s = warp * D + slot;
      unsigned char* stg = stages + (size_t)s * SH::STAGE;
      const uint32_t sha = (uint32_t)(reinterpret_cast<uint64_t>(a_data + (pa - 1)) & 15ull);
      const uint32_t shb = (uint32_t)(reinterpret_cast<uint64_t>(b_data + (pb - 1)) & 15ull);
      const double* __restrict__ As = reinterpret_cast<const double*>(stg + sha);
      const double* __restrict__ Bs = reinterpret_cast<const double*>(stg + SH::ABUF + shb);
      if (TRACE) {
        if (rec != nullptr && i < TRACE_ENTRIES) rec[4 + 4 * i] = clock_now();
      }
      mbar_wait(&full[s], ph);
      if (STAG) {
        const bool stag_me = NC > 4 ? (((warp >> 2) & 1) != 0) : (blockIdx.x * 2 >= gridDim.x);
        if (i == 0 && stag_me) {
          const unsigned long long t0 = clock_now();
          while (clock_now() - t0 < 1000ull) {
          }
        }
      }
      if (TRACE) {
        if (rec != nullptr && i < TRACE_ENTRIES) {
          rec[4 + 4 * i + 1] = clock_now();
          rec[4 + 4 * i + 3] = (unsigned long long)pc;
        }
      }
      mma_entry<M, N, K>(As, Bs, acc, g, t, sha, shb);
      __syncwarp();  // every lane has read its fragments: the stage can be handed back to the producer
      if (FLUSH == 2) {
        if (pending >= 0) {  // the bulk reduction issued one entry ago has read its stage by now
          if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(&empty[pending]);
          }
          pending = -1;
        }
        const bool run_ends = (i + 1 == len) || (pe[5] != pc);
        if (run_ends) {
          unsigned long long tf0 = 0;
          if (TRACE) tf0 = clock_now();
          flush_acc_bulk<M, N, K>(c_data, pc, acc, g, t, lane, stg);
          pending = s;
          if (TRACE) {
            t_flush += clock_now() - tf0;
            ++n_flush;
          }
        }
        else if (lane == 0) {
          mbar_arrive(&empty[s]);
        }
      }
      else {
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      if (TRACE) {
        if (rec != nullptr && i < TRACE_ENTRIES) rec[4 + 4 * i + 2] = clock_now();
      }
      if (++slot == D) {
        slot = 0;
        ph ^= 1u;
      }
    }
    if (FLUSH == 2) {
      // every bulk reduction of this warp is performed before it exits (the stage memory is released with the CTA)
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __syncwarp();
    }
    else if (cur_c >= 0) {
      flush_acc<M, N, K, HINT>(c_data, cur_c, acc, g, t, pol_c);
    }
    if (TRACE) {
      if (rec != nullptr) {
        rec[124] = n_flush + (FLUSH == 0 ? 1 : 0);
        rec[125] = t_flush;
        rec[126] = clock_now();
        rec[127] = globaltimer_now();
      }
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace smm
