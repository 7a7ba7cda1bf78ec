// ftsgemm_kernel.cuh -- the fused online-ABFT SGEMM kernel for sm_100a (B200).
//
// Replaces the reference's code-generated CUDA-core kernels
//   sgemm_{small..huge}    (/root/reference/kernel/ft_sgemm/include_code_gen/sgemm_*.cuh:11)     [FT = false]
//   ft_sgemm_{small..huge} (/root/reference/kernel/ft_sgemm/include_code_gen/ft_sgemm_*.cuh:11)  [FT = true]
// with ONE warp-specialised, persistent tcgen05 kernel template:
//
//   warp 0   TMA producer   : cp.async.bulk.tensor 32(M|N) x 32(K) fp32 boxes, 128B swizzle with 32B atoms,
//                             into a STAGES-deep shared-memory ring (full/empty mbarriers)
//   warp 1   MMA issuer     : one thread issues tcgen05.mma.kind::tf32 (UMMA 128 x BN x 8, FP32 accumulate in TMEM);
//                             when FT, a second UMMA 128 x 16 x 8 per k-step multiplies the same A tile by the
//                             *checksum columns* of the B tile, so the expected row checksums accumulate in 16 extra
//                             TMEM columns next to the data ("checksum GEMM rides the same TMEM tile")
//   warp 2   TMEM allocator
//   warps 4-7 epilogue      : tcgen05.ld the accumulator (lane = row), per-row detect / locate / correct against the
//                             checksum columns, then C = alpha*acc + beta*C with coalesced column-major stores
//
// ABFT scheme.  For a CTA tile with rows I (128) and columns J (BN), with b~ the TF32 value
// the tensor core actually consumes:
//   encode    (pre-pass, encode.cuh; reference ft_sgemm_huge.cuh:150-168)
//             e[k] = sum_{n in J} b~[n,k]         w[k] = sum_{n in J} (n-n0+1) b~[n,k]      (3-way TF32 split each)
//   checksum GEMM (tensor core; reference :171-213)
//             r1[m] = sum_k a~[m,k] e[k]          r2[m] = sum_k a~[m,k] w[k]
//   detect    (epilogue; reference :328-421)
//             d1[m] = r1[m] - sum_n acc[m,n]      d2[m] = r2[m] - sum_n (n-n0+1) acc[m,n]
//             flagged iff |d1| > tau_abs + tau_rel * sum_n |acc[m,n]|
//   locate    column j = round(d2/d1) - 1   (weighted checksum; the reference intersects a row and a column residual)
//   correct   acc[m,j] = r1[m] - sum_{n != j} acc[m,n]  (recomputed, so it also repairs Inf/NaN/huge upsets;
//             reference :422-485 adds the row residual)
// One error per (row, tile) is correctable, i.e. up to 128 per tile (the reference: one per tile per check).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "ptx.cuh"

namespace ftsgemm {

constexpr int kBM = 128;          // UMMA M (cta_group::1)
constexpr int kBK = 32;           // K extent of one shared-memory stage (4 UMMA k-steps of 8)
constexpr int kAtomMN = 32;       // floats per 128-byte swizzle row
constexpr int kChkCols = 16;      // UMMA N of the checksum GEMM (6 columns used)
constexpr int kChkUsed = 6;
constexpr int kThreads = 256;
constexpr int kMaxFaults = 8;
constexpr int kMaxEvents = 16;

struct DeviceFault {
  int row, col, mode;
  float add_value;
  unsigned int xor_mask;
};
struct DeviceEvent {
  int row, col;
  float residual, corrected_value;
  int status;
};
struct DeviceStats {
  unsigned long long tiles, rows_checked, detected, corrected, uncorrectable, checksum_faults;
  unsigned int max_abs_bits, max_rel_bits;
  int n_events;
  DeviceEvent events[kMaxEvents];
};

struct KernelParams {
  int M, N, K;
  float *C;
  int ldc;
  float alpha, beta;
  // tile schedule
  int tiles_m, tiles_n, group_n;
  // UMMA shared-memory descriptor parameters (runtime so the bring-up probe can sweep them)
  unsigned int lbo_bytes, sbo_bytes, layout_type, kstep_bytes;
  // fault tolerance
  float tau_abs, tau_rel;
  int detect_only;
  int inject_mode;
  float selftest_value;
  int selftest_row, selftest_col;
  int n_faults;
  DeviceFault faults[kMaxFaults];
  DeviceStats *stats;
};

template <int BN, bool FT>
struct TileCfg {
  static constexpr int kABytes = kBM * kBK * 4;
  static constexpr int kBBytes = BN * kBK * 4;
  static constexpr int kCBytes = FT ? kAtomMN * kBK * 4 : 0;
  static constexpr int kStageBytes = kABytes + kBBytes + kCBytes;
  static constexpr int kAccStride = FT ? BN + 32 : BN;            // TMEM columns per accumulator stage
  static constexpr int kAccStages = (2 * kAccStride <= 512) ? 2 : 1;
  static constexpr int kTmemNeeded = kAccStages * kAccStride;
  static constexpr int kTmemCols = kTmemNeeded <= 32 ? 32 : kTmemNeeded <= 64 ? 64 : kTmemNeeded <= 128 ? 128
                                   : kTmemNeeded <= 256 ? 256 : 512;
  static constexpr int kMaxSmem = 227 * 1024 - 1024 /*alignment slack*/ - 256 /*barriers*/;
  static constexpr int kStagesFit = kMaxSmem / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }

__device__ __forceinline__ void decode_tile(const KernelParams &p, int t, int &m_blk, int &n_blk) {
  const int per_group = p.group_n * p.tiles_m;
  const int g = t / per_group;
  const int first_n = g * p.group_n;
  const int gsz = min(p.group_n, p.tiles_n - first_n);
  const int local = t - g * per_group;
  n_blk = first_n + local % gsz;
  m_blk = local / gsz;
}

template <int BN, bool FT>
__global__ void __launch_bounds__(kThreads, 1)
ftsgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmChk, const KernelParams p) {
  using Cfg = TileCfg<BN, FT>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kAccStages = Cfg::kAccStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B-swizzle atoms need 1024B alignment
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t *tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
    if (FT) ptx::tma_prefetch_desc(&tmChk);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0 && lane == 0) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_blk, n_blk;
      decode_tile(p, t, m_blk, n_blk);
      const int m0 = m_blk * kBM, n0 = n_blk * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        ptx::mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
        const uint32_t sA = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sB = sA + Cfg::kABytes;
        const int k0 = kb * kBK;
#pragma unroll
        for (int i = 0; i < kBM / kAtomMN; ++i)
          ptx::tma_load_2d(sA + i * (kBK * 128), &tmA, full_bar(stage), m0 + i * kAtomMN, k0);
#pragma unroll
        for (int i = 0; i < BN / kAtomMN; ++i)
          ptx::tma_load_2d(sB + i * (kBK * 128), &tmB, full_bar(stage), n0 + i * kAtomMN, k0);
        if (FT) ptx::tma_load_2d(sB + Cfg::kBBytes, &tmChk, full_bar(stage), n_blk * kAtomMN, k0);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================================================================== MMA issuer
    const uint32_t idesc_main = ptx::make_idesc_tf32(kBM, BN, 1, 1);
    const uint32_t idesc_chk = ptx::make_idesc_tf32(kBM, kChkCols, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_main = tmem_base + acc * Cfg::kAccStride;
      const uint32_t d_chk = d_main + BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sA = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sB = sA + Cfg::kABytes;
        const uint32_t sC = sB + Cfg::kBBytes;
#pragma unroll
        for (int j = 0; j < kBK / 8; ++j) {
          const uint64_t da = ptx::make_smem_desc(sA + j * p.kstep_bytes, p.lbo_bytes, p.sbo_bytes, p.layout_type);
          const uint64_t db = ptx::make_smem_desc(sB + j * p.kstep_bytes, p.lbo_bytes, p.sbo_bytes, p.layout_type);
          const uint32_t accum = (kb | j) != 0 ? 1u : 0u;
          ptx::mma_tf32(d_main, da, db, idesc_main, accum);
          if (FT) {
            const uint64_t dc = ptx::make_smem_desc(sC + j * p.kstep_bytes, p.lbo_bytes, p.sbo_bytes, p.layout_type);
            ptx::mma_tf32(d_chk, da, dc, idesc_chk, accum);
          }
        }
        ptx::mma_commit(empty_bar(stage));  // frees the smem slot once these MMAs have read it
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      ptx::mma_commit(tfull_bar(acc));  // accumulator (data + checksum columns) complete
      if (kAccStages == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      } else {
        acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue (4 warps, lane = row)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_blk, n_blk;
      decode_tile(p, t, m_blk, n_blk);
      const int m0 = m_blk * kBM, n0 = n_blk * BN;
      const int m = m0 + row;
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::kAccStride;

      int fix_col = -1;       // tile-local column whose value is replaced in the store pass
      float fix_val = 0.0f;

      if (FT) {
        // ---- fault injection into the TMEM accumulator (reference: ft_sgemm_huge.cuh:324-327) ----
        if (p.inject_mode == 1) {
          if ((p.selftest_row >> 5) == q && p.selftest_col < BN) {
            uint32_t x = ptx::tmem_ld_x1(taddr + p.selftest_col);
            ptx::tmem_wait_ld();
            if (lane == (p.selftest_row & 31)) x = f2u(u2f(x) + p.selftest_value);
            ptx::tmem_st_x1(taddr + p.selftest_col, x);
            ptx::tmem_wait_st();
          }
        } else if (p.inject_mode == 2) {
          for (int f = 0; f < p.n_faults; ++f) {
            const int tr = p.faults[f].row - m0, tc = p.faults[f].col - n0;
            if (tr >= 0 && tr < kBM && tc >= 0 && tc < BN && (tr >> 5) == q) {  // warp-uniform
              uint32_t x = ptx::tmem_ld_x1(taddr + tc);
              ptx::tmem_wait_ld();
              if (lane == (tr & 31))
                x = p.faults[f].mode == 0 ? f2u(u2f(x) + p.faults[f].add_value) : (x ^ p.faults[f].xor_mask);
              ptx::tmem_st_x1(taddr + tc, x);
              ptx::tmem_wait_st();
            }
          }
        }
        // ---- pass 1: actual row checksums (thread-local: lane == row) ----
        float s1 = 0.0f, s2 = 0.0f, sabs = 0.0f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_x32(taddr + c * 32, v);
          ptx::tmem_wait_ld();
          const float wbase = static_cast<float>(c * 32 + 1);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float f = u2f(v[i]);
            s1 += f;
            s2 = fmaf(f, wbase + static_cast<float>(i), s2);
            sabs += fabsf(f);
          }
        }
        uint32_t e[8];
        ptx::tmem_ld_x8(taddr + BN, e);
        ptx::tmem_wait_ld();
        const float r1 = u2f(e[0]) + (u2f(e[1]) + u2f(e[2]));
        const float r2 = u2f(e[3]) + (u2f(e[4]) + u2f(e[5]));
        const float d1 = r1 - s1, d2 = r2 - s2;
        const float thr = p.tau_abs + p.tau_rel * sabs;
        const bool flagged = !(fabsf(d1) <= thr);  // also true for NaN
        const unsigned flag_mask = __ballot_sync(0xffffffffu, flagged);

        // fault-free residual statistics (threshold calibration)
        {
          float ra = flagged ? 0.0f : fabsf(d1);
          float rr = flagged ? 0.0f : fabsf(d1) / fmaxf(sabs, 1e-30f);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            ra = fmaxf(ra, __shfl_xor_sync(0xffffffffu, ra, o));
            rr = fmaxf(rr, __shfl_xor_sync(0xffffffffu, rr, o));
          }
          if (lane == 0 && p.stats) {
            atomicMax(&p.stats->max_abs_bits, f2u(ra));
            atomicMax(&p.stats->max_rel_bits, f2u(rr));
            atomicAdd(&p.stats->rows_checked, 32ull);
            if (q == 0) atomicAdd(&p.stats->tiles, 1ull);
          }
        }

        if (flag_mask != 0u) {  // rare slow path, warp-uniform
          // candidate column from the weighted checksum; Inf/NaN rows fall back to the largest-magnitude element
          int j = -1;
          bool use_argmax = false;
          if (flagged) {
            if (isfinite(d1) && isfinite(d2) && d1 != 0.0f) {
              const float jf = d2 / d1;
              const float jr = rintf(jf);
              if (fabsf(jf) < 0.5f) j = -2;  // d2 ~ 0: the checksum column itself was hit, data are intact
              else if (jr >= 1.0f && jr <= static_cast<float>(BN) && fabsf(jf - jr) <= 0.3f) j = static_cast<int>(jr) - 1;
              else j = -1;
            } else {
              use_argmax = true;
            }
          }
          const unsigned argmax_mask = __ballot_sync(0xffffffffu, use_argmax);
          if (argmax_mask != 0u) {
            float best = -1.0f;
            int bj = 0;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
              uint32_t v[32];
              ptx::tmem_ld_x32(taddr + c * 32, v);
              ptx::tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float a = fabsf(u2f(v[i]));
                const bool better = !(a <= best);  // NaN wins
                if (better) {
                  best = (a == a) ? a : CUDART_INF_F;
                  bj = c * 32 + i;
                }
              }
            }
            if (use_argmax) j = bj;
          }
          // recompute the row checksums without column j
          float x1 = 0.0f, x2 = 0.0f;
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_x32(taddr + c * 32, v);
            ptx::tmem_wait_ld();
            const float wbase = static_cast<float>(c * 32 + 1);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float f = (c * 32 + i == j) ? 0.0f : u2f(v[i]);
              x1 += f;
              x2 = fmaf(f, wbase + static_cast<float>(i), x2);
            }
          }
          if (flagged) {
            int status;
            float vc = 0.0f;
            if (j == -2) {
              status = 4;
            } else if (j < 0) {
              status = 3;
            } else {
              vc = r1 - x1;
              const float wj = static_cast<float>(j + 1);
              const float e2 = (r2 - x2) - wj * vc;  // second checksum must agree with a single error at j
              const float tol = static_cast<float>(BN) * (p.tau_abs + p.tau_rel * (sabs == sabs && isfinite(sabs) ? sabs : fabsf(x1) + fabsf(vc)));
              status = (fabsf(e2) <= tol) ? 1 : 3;
            }
            if (status == 1 && p.detect_only) status = 2;
            if (status == 1) {
              fix_col = j;
              fix_val = vc;
            }
            if (p.stats) {
              atomicAdd(&p.stats->detected, 1ull);
              if (status == 1) atomicAdd(&p.stats->corrected, 1ull);
              if (status == 3) atomicAdd(&p.stats->uncorrectable, 1ull);
              if (status == 4) atomicAdd(&p.stats->checksum_faults, 1ull);
              const int slot = atomicAdd(&p.stats->n_events, 1);
              if (slot < kMaxEvents) {
                DeviceEvent ev;
                ev.row = m;
                ev.col = j >= 0 ? n0 + j : -1;
                ev.residual = d1;
                ev.corrected_value = vc;
                ev.status = status;
                p.stats->events[slot] = ev;
              }
            }
          }
        }
      }

      // ---- store pass: C = alpha*acc + beta*C, column-major (lane = consecutive m -> 128B coalesced) ----
      const bool row_ok = m < p.M;
      float *crow = p.C + m;
      const bool full_n = (n0 + BN <= p.N);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_x32(taddr + c * 32, v);
        ptx::tmem_wait_ld();
        if (FT && (fix_col >> 5) == c && fix_col >= 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i == (fix_col & 31)) v[i] = f2u(fix_val);
        }
        const int nb = n0 + c * 32;
        if (row_ok) {
          if (full_n) {
            if (p.beta == 0.0f) {
#pragma unroll
              for (int i = 0; i < 32; ++i) crow[static_cast<size_t>(nb + i) * p.ldc] = p.alpha * u2f(v[i]);
            } else {
              float old[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) old[i] = crow[static_cast<size_t>(nb + i) * p.ldc];
#pragma unroll
              for (int i = 0; i < 32; ++i)
                crow[static_cast<size_t>(nb + i) * p.ldc] = p.alpha * u2f(v[i]) + p.beta * old[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (nb + i < p.N) {
                float *dst = crow + static_cast<size_t>(nb + i) * p.ldc;
                const float o = (p.beta == 0.0f) ? 0.0f : p.beta * (*dst);
                *dst = p.alpha * u2f(v[i]) + o;
              }
            }
          }
        }
      }
      // release this accumulator stage back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
      if (kAccStages == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      } else {
        acc_phase ^= 1u;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Encode pre-pass (reference: ft_sgemm_huge.cuh:150-168 ENCODE of B, done there per CTA and per k-step with
// shuffles; here once per GEMM and per BN-wide column block, because a CUDA-core re-read of every shared-memory
// stage does not fit next to a tensor-core main loop -- DESIGN.md section 3).
//   chk[k][t*32 + 0..2] = 3-way TF32 split of  e = sum_{n in block t} tf32(B[n,k])
//   chk[k][t*32 + 3..5] = 3-way TF32 split of  w = sum_{n in block t} (n - n0 + 1) * tf32(B[n,k])
// Sums are accumulated in FP64, so the three TF32 terms carry the checksum to ~2^-33 relative.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_bits(float x, int rounding) {
  uint32_t u = __float_as_uint(x);
  if (rounding == 1) u += 0x1000u;  // round-to-nearest (ties away), like cvt.rna.tf32.f32
  if (rounding != 2) u &= 0xFFFFE000u;
  return __uint_as_float(u);
}
__device__ __forceinline__ void split3_tf32(double x, float &h, float &m, float &l) {
  h = tf32_bits(static_cast<float>(x), 0);
  double r = x - static_cast<double>(h);
  m = tf32_bits(static_cast<float>(r), 0);
  r -= static_cast<double>(m);
  l = tf32_bits(static_cast<float>(r), 0);
}

constexpr int kEncWarps = 8;
constexpr int kEncKPerWarp = 4;

__global__ void __launch_bounds__(kEncWarps * 32)
encode_b_kernel(const float *__restrict__ B, int N, int K, int ldb, int BN, float *__restrict__ chk, int chk_ld,
                int rounding) {
  const int t = blockIdx.x;
  const int n0 = t * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kbase = (blockIdx.y * kEncWarps + warp) * kEncKPerWarp;
  double e[kEncKPerWarp], w[kEncKPerWarp];
#pragma unroll
  for (int u = 0; u < kEncKPerWarp; ++u) e[u] = w[u] = 0.0;
  for (int j = lane; j < BN; j += 32) {
    const int n = n0 + j;
    if (n < N) {
#pragma unroll
      for (int u = 0; u < kEncKPerWarp; ++u) {
        const int k = kbase + u;
        if (k < K) {
          const float b = tf32_bits(__ldg(B + static_cast<size_t>(k) * ldb + n), rounding);
          e[u] += static_cast<double>(b);
          w[u] += static_cast<double>(b) * static_cast<double>(j + 1);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kEncKPerWarp; ++u) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e[u] += __shfl_xor_sync(0xffffffffu, e[u], o);
      w[u] += __shfl_xor_sync(0xffffffffu, w[u], o);
    }
    const int k = kbase + u;
    if (k < K) {  // every lane holds the totals; lane i writes column i of the 32-float block (6 used, rest zero)
      float eh, em, el, wh, wm, wl;
      split3_tf32(e[u], eh, em, el);
      split3_tf32(w[u], wh, wm, wl);
      const float val = lane == 0 ? eh : lane == 1 ? em : lane == 2 ? el : lane == 3 ? wh : lane == 4 ? wm
                        : lane == 5 ? wl : 0.0f;
      chk[static_cast<size_t>(k) * chk_ld + t * kAtomMN + lane] = val;
    }
  }
}

}  // namespace ftsgemm
