// tc_gemm.cu -- the path's dense contraction on Blackwell tensor cores (tcgen05 + TMEM + TMA).
//
//   D[M x N] = alpha * A[M x K] . B[N x K]^T (+ bias[N])      A, B bf16 (K-major), fp32 accumulate
//
// Every GEMM-shaped product of the hot path at hidden >= 128 goes through this kernel:
//   GIN MLP Linear1 / Linear2          (gcc/models/gin.py:107-116; z = x W^T + b, BatchNorm statistics
//                                       of z fused into the epilogue -- gin.py:115 needs the column
//                                       mean / variance over all N rows)
//   their input gradients dX = dZ . W  (B operand = W^T kept as a second bf16 copy)
//   weight gradients dW = dZ^T . X     (operands transposed once, split-K over the rows)
//   MoCo logits q . queue^T            (gcc/contrastive/memory_moco.py:33-44) and dq = P . queue
// Design (one CTA per SM, persistent over 128-row output tiles):
//   warp 0    TMA producer: cp.async.bulk.tensor 2-D boxes (128 x 64 of A, BN x 64 of B, 128-byte
//             swizzle) into a 4-stage shared-memory ring, completion on mbarriers;
//   warp 1    allocates TMEM and issues tcgen05.mma.cta_group::1.kind::f16 (M 128, N = BN, K 16) from
//             one elected lane; tcgen05.commit releases the ring slot / publishes the accumulator;
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per warp), bias / scale, fp32 and/or bf16
//             stores, column sums and sums of squares (float64 atomics, like the SIMT kernels).
//   Two accumulator stages in TMEM (2 x BN <= 512 columns): the epilogue of tile t overlaps the
//   MMAs of tile t+1.
// Shared-memory operand layout = the canonical K-major SWIZZLE_128B layout (8-row groups 1024 B apart),
// which is exactly what a TMA box with CU_TENSOR_MAP_SWIZZLE_128B writes.
#include "tc_gemm.cuh"

#ifndef GCCB_EMU
#include <cuda.h>
#include <cuda_bf16.h>

namespace gccb {
namespace tc {

constexpr int BM = 128, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int BIAS_SMEM = 512;                             // floats of bias staged in shared memory
// Epilogue warps: a warp may only touch the TMEM lane group (warp % 4), so the 128 accumulator rows always need
// 4 warps; wide tiles are split by COLUMNS over PARTS such sets (BN = 256: 16 warps of 64 columns each).  One
// epilogue warp per scheduler issues an instruction every ~5 cycles (ncu: 18 % issue-slot use, the tile loop was
// bound by this warp's dependent chain, not by TMA or the tensor pipe); four per scheduler hide each other.
template <int BN> struct Epi {
  static constexpr int PARTS = BN >= 256 ? 4 : BN >= 128 ? 2 : 1;
  static constexpr int WARPS = 4 * PARTS, THREADS_EPI = 32 * WARPS, THREADS = 64 + THREADS_EPI;
  static constexpr int COLS = BN / PARTS;                  // columns per warp (a multiple of 32)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Spin on a phase parity.  A protocol bug would otherwise hang the GPU until the driver's watchdog:
// after ~2 s of spinning the kernel traps (the launch then fails loudly instead).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint64_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1ull << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (unused for swizzled K-major) = 1 in [16,30), stride byte offset
// (8-row group pitch, 1024 B) >> 4 in [32,46), version 1 in [46,48), layout type 2 = SWIZZLE_128B in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1 << 4), a/b_format BF16 (1 << 7, 1 << 10), K-major both
// (bits 15, 16 = 0), n_dim = N >> 3 at bit 17, m_dim = M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct GemmArgs {
  int M_cap, N, K, ldo;
  const int32_t* m_dev;       // optional: number of valid rows (device)
  const float* bias;          // [N] or null
  float alpha;
  float* out_f32;             // [M_cap][ldo] (or [splits][M_cap][ldo] partials when splits > 1) or null
  __nv_bfloat16* out_bf16;    // [M_cap][ldo] or null
  double* colstats;           // [2][N] (sum, sum of squares of the stored values) or null
  int splits;
};

// BRES: the whole B operand (N == BN rows, K <= 256) stays resident in shared memory for the life of the CTA --
// the GIN MLP case (a 256 x 256 weight matrix = 128 KB of bf16): only the A tiles stream through the ring, so a
// 128-row tile costs 64 KB of L2 -> shared-memory traffic instead of 192 KB.
template <int BN, bool BRES>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = BRES ? A_BYTES : A_BYTES + B_BYTES;
  static constexpr int BRES_KB = 4;                        // k-blocks of the resident operand (K <= 256)
  static constexpr size_t BRES_BYTES = BRES ? (size_t)BRES_KB * B_BYTES : 0;
  static constexpr int STAGE_F = 128 * 33;                 // epilogue transpose buffer (floats)
  static constexpr size_t TOTAL = 1024 + BRES_BYTES + (size_t)STAGES * STAGE_BYTES + STAGE_F * 4 + 2 * 256 * 4 + 256;
};

template <int BN, bool BRES>
__global__ void __launch_bounds__(Epi<BN>::THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  using S = Smem<BN, BRES>;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* bres = base;                                        // [BRES_KB][B] (BRES only)
  unsigned char* tiles = base + S::BRES_BYTES;                       // [STAGES][A | B]  (BRES: [STAGES][A])
  float* stage_t = (float*)(tiles + (size_t)STAGES * S::STAGE_BYTES);   // [128][33]
  float* colpart = stage_t + S::STAGE_F;                              // [512]: bias staging
  uint64_t* bars = (uint64_t*)(colpart + 2 * 256);
  uint64_t* full = bars;                 // [STAGES]   TMA -> MMA
  uint64_t* empty = bars + STAGES;       // [STAGES]   MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;   // [2]        MMA -> epilogue
  uint64_t* tempty = bars + 2 * STAGES + 2;   // [2]   epilogue -> MMA
  uint64_t* bfull = bars + 2 * STAGES + 4;    // [1]   resident B operand landed
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = g.m_dev ? min(max(*g.m_dev, 0), g.M_cap) : g.M_cap;
  const int m_tiles = (M + BM - 1) / BM, n_tiles = g.N / BN;
  const int kb_total = g.K / BK;
  const int split = blockIdx.y;
  const int kb_per = (kb_total + g.splits - 1) / g.splits;
  const int kb0 = split * kb_per, kb1 = min(kb_total, kb0 + kb_per);
  const int nkb = max(kb1 - kb0, 0);
  const int num_tiles = m_tiles * n_tiles;
  constexpr uint32_t TMEM_COLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], Epi<BN>::WARPS); }
    mbar_init(bfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                       // TMEM allocation: one full warp; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // bias (zeros when absent) in shared memory when it fits: the epilogue then needs no global loads
  float* bias_s = colpart;                                            // [BIAS_SMEM]
  const bool bias_in_smem = g.N <= BIAS_SMEM;
  if (bias_in_smem)
    for (int i = threadIdx.x; i < g.N; i += blockDim.x) bias_s[i] = g.bias ? g.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nkb > 0) {
    if (warp == 0) {
      // ===== TMA producer ============================================================================
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        int stage = 0;
        uint32_t phase = 0;
        if (BRES && blockIdx.x < num_tiles) {              // the whole B operand, once
          mbar_expect_tx(bfull, (uint32_t)(nkb * S::B_BYTES));
          for (int kb = 0; kb < nkb; ++kb) tma_load_2d(bres + (size_t)kb * S::B_BYTES, &map_b, kb * BK, 0, bfull);
        }
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            unsigned char* sa = tiles + (size_t)stage * S::STAGE_BYTES;
            mbar_expect_tx(&full[stage], S::STAGE_BYTES);
            tma_load_2d(sa, &map_a, kb * BK, mt * BM, &full[stage]);
            if (!BRES) tma_load_2d(sa + S::A_BYTES, &map_b, kb * BK, nt * BN, &full[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer (one elected lane) ===========================================================
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc(BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        if (BRES && blockIdx.x < num_tiles) mbar_wait(bfull, 0);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          mbar_wait(&tempty[acc], acc_phase ^ 1);          // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(tiles + (size_t)stage * S::STAGE_BYTES);
            const uint32_t sb = BRES ? smem_u32(bres + (size_t)kb * S::B_BYTES) : sa + S::A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advancing 16 elements (32 B) along K inside the 128-byte swizzle atom = +32 B on the start address
              umma_bf16(d_tmem, umma_desc(sa + k * UMMA_K * 2), umma_desc(sb + k * UMMA_K * 2), idesc,
                        (uint32_t)((kb | k) != 0));
            }
            umma_commit(&empty[stage]);                    // ring slot free once these MMAs have read it
            if (kb == nkb - 1) umma_commit(&tfull[acc]);   // accumulator complete
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    } else {
      // ===== epilogue warps 2..: TMEM lanes 32 * (warp % 4) .. +31, columns [part * COLS, (part + 1) * COLS) ====
      const int ew = warp & 3;                              // TMEM lane group this warp may access
      const int part = (warp - 2) >> 2;
      constexpr int COLS = Epi<BN>::COLS;
      const int et = threadIdx.x - 64;                      // 0 .. THREADS_EPI - 1
      const int row_in_tile = ew * 32 + lane;
      int acc = 0;
      uint32_t acc_phase = 0;
      float* outf = g.out_f32 ? g.out_f32 + (size_t)split * g.M_cap * g.ldo : nullptr;
      const bool wide_f32 = outf && ((((uintptr_t)outf) | ((size_t)g.ldo * 4)) & 31) == 0;
      const bool wide_bf16 = g.out_bf16 && ((((uintptr_t)g.out_bf16) | ((size_t)g.ldo * 2)) & 31) == 0;
      if (g.colstats) {                                     // per-warp column accumulators (n_tiles == 1 with statistics)
        for (int i = lane; i < COLS; i += 32) {               // own columns only: the warps of a lane group share a row
          stage_t[ew * (2 * BN) + part * COLS + i] = 0.f;
          stage_t[ew * (2 * BN) + BN + part * COLS + i] = 0.f;
        }
        __syncwarp();
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
        const int row = mt * BM + row_in_tile;
        const bool row_ok = row < M;
        const bool full_tile = (mt + 1) * BM <= M;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = part * COLS; c0 < (part + 1) * COLS; c0 += 32) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int col0 = nt * BN + c0;
          float v[32];
          if (bias_in_smem) {                               // bias (or zeros) staged in shared memory: 8 x LDS.128
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = b4[j];
              v[4 * j] = fmaf(__uint_as_float(r[4 * j]), g.alpha, b.x);
              v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), g.alpha, b.y);
              v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), g.alpha, b.z);
              v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), g.alpha, b.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), g.alpha, g.bias ? g.bias[col0 + j] : 0.f);
          }
          if (!full_tile) {                                 // only the last row tile has rows to mask
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = row_ok ? v[j] : 0.f;
          }
          if (row_ok) {
            // a TMEM lane holds one ROW: each lane stores its 32 columns itself.  256-bit stores (sm_100) when the
            // row pitch allows: every store fills whole 32-byte sectors, half the store instructions
            if (outf) {
              float* dst = outf + (size_t)row * g.ldo + col0;
              if (wide_f32) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * j),
                               "f"(v[8 * j]), "f"(v[8 * j + 1]), "f"(v[8 * j + 2]), "f"(v[8 * j + 3]), "f"(v[8 * j + 4]),
                               "f"(v[8 * j + 5]), "f"(v[8 * j + 6]), "f"(v[8 * j + 7])
                               : "memory");
              } else {
                float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
                for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
            }
            if (g.out_bf16) {
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                pk[j] = *reinterpret_cast<uint32_t*>(&p2);
              }
              __nv_bfloat16* dst = g.out_bf16 + (size_t)row * g.ldo + col0;
              if (wide_bf16) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16 * j),
                               "r"(pk[8 * j]), "r"(pk[8 * j + 1]), "r"(pk[8 * j + 2]), "r"(pk[8 * j + 3]), "r"(pk[8 * j + 4]),
                               "r"(pk[8 * j + 5]), "r"(pk[8 * j + 6]), "r"(pk[8 * j + 7])
                               : "memory");
              } else {
                uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
                for (int j = 0; j < 4; ++j) d4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              }
            }
          }
          if (g.colstats) {
            // column sums over the warp's 32 rows without shared memory or barriers: a shuffle transpose-reduce
            // (16 + 8 + 4 + 2 + 1 exchanges per quantity) leaves column `lane` of the chunk on lane `lane`; the
            // warp then adds it to its private accumulator row (flushed once per CTA, below)
            float q[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) q[j] = v[j] * v[j];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < off; ++i) {
                const float sv = upper ? v[i] : v[i + off], kv = upper ? v[i + off] : v[i];
                v[i] = kv + __shfl_xor_sync(0xffffffffu, sv, off);
                const float sq = upper ? q[i] : q[i + off], kq = upper ? q[i + off] : q[i];
                q[i] = kq + __shfl_xor_sync(0xffffffffu, sq, off);
              }
            }
            float* accw = stage_t + ew * (2 * BN);          // [4 warps][2][BN] floats (BN <= 256: 8 KB of the 16.5 KB)
            accw[c0 + lane] += v[0];
            accw[BN + c0 + lane] += q[0];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (g.colstats) {                                     // one float64 atomic per column and CTA
        asm volatile("bar.sync 1, %0;" ::"n"(Epi<BN>::THREADS_EPI) : "memory");
        for (int i = et; i < 2 * BN; i += Epi<BN>::THREADS_EPI) {
          const float t = stage_t[i] + stage_t[2 * BN + i] + stage_t[4 * BN + i] + stage_t[6 * BN + i];
          const int which = i / BN, c = i - which * BN;
          if (blockIdx.x < num_tiles) atomicAdd(&g.colstats[(size_t)which * g.N + c], (double)t);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// out[r][c] = alpha * sum_s part[s][r][c] (+ bias[c]) (+ beta * out[r][c]) for c < n_out; optional bf16 copy.
// Partials have row pitch ldp, the output row pitch ldo.  Fixed summation order: deterministic.
__global__ void __launch_bounds__(256)
tc_splitk_reduce_kernel(const float* __restrict__ part, int splits, int M_cap, int n_out, int ldp, int ldo,
                        const int32_t* __restrict__ m_dev, float alpha, float beta, const float* __restrict__ bias,
                        float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
  const int M = m_dev ? min(max(*m_dev, 0), M_cap) : M_cap;
  const size_t total = (size_t)M * n_out;
  const size_t stride = (size_t)M_cap * ldp;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / n_out), c = (int)(idx - (size_t)r * n_out);
    const size_t ip = (size_t)r * ldp + c, o = (size_t)r * ldo + c;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += part[(size_t)sp * stride + ip];
    s *= alpha;
    if (bias) s += bias[c];
    if (out_f32) {
      if (beta != 0.f) s = fmaf(beta, out_f32[o], s);
      out_f32[o] = s;
    }
    if (out_bf16) out_bf16[o] = __float2bfloat16_rn(s);
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 row-major [rows][cols] -> boxes of [box_rows][64], 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* ptr, int rows, int cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_last_error("tc_gemm: cuTensorMapEncodeTiled is not available"); return GCCB_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("tc_gemm: cuTensorMapEncodeTiled failed (%d)", (int)r); return GCCB_ERR_CUDA; }
  return GCCB_OK;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, bool BRES = false>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const GemmArgs& g, cudaStream_t stream) {
  auto k = tc_gemm_kernel<BN, BRES>;
  const size_t smem = Smem<BN, BRES>::TOTAL;
  gccb::ensure_dyn_smem(k, smem);
  const int m_tiles = (g.M_cap + BM - 1) / BM, tiles = m_tiles * (g.N / BN);
  int gx = sm_count() / g.splits;
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  dim3 grid(gx, g.splits);
  ++gccb::g_launch_count;
  k<<<grid, Epi<BN>::THREADS, smem, stream>>>(ma, mb, g);
  return GCCB_OK;
}

// Internal entry used by the GIN / MoCo tensor-core paths and by gccb_tc_gemm_bf16.
// splits > 1 (or an accumulating / narrow / oddly pitched output): `scratch` must hold splits * M_cap * N floats; the partial sums are reduced in a fixed order.
int gemm_bf16(const void* A, const void* B, int M_cap, int N, int K, const int32_t* m_dev, const float* bias,
              float alpha, float* out_f32, void* out_bf16, int ldo, double* colstats, int splits, float* scratch,
              cudaStream_t stream, float beta, int n_out) {
  if (n_out < 0) n_out = N;
  if (!A || !B || M_cap <= 0 || N <= 0 || K <= 0 || (K % BK) || (N % 32) || n_out > N || ldo < n_out || splits < 1 ||
      (!out_f32 && !out_bf16)) {
    set_last_error("tc_gemm: bad argument (need K %% 64 == 0, N %% 32 == 0, ldo >= n_out)");
    return GCCB_ERR_BADARG;
  }
  const int BN = (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : (N % 64 == 0) ? 64 : 32;
  if (splits > K / BK) splits = K / BK;
  {                                         // no empty split: every partial buffer is written
    const int kb_total = K / BK, per = (kb_total + splits - 1) / splits;
    splits = (kb_total + per - 1) / per;
  }
  // partial sums go through `scratch` and the reduce kernel when K is split, and also when the output is not
  // a plain [M][N] store (accumulate into it, narrower than N, row pitch the epilogue's 16-byte stores cannot use)
  const bool via_scratch = splits > 1 || beta != 0.f || n_out != N || (ldo % 8) != 0;
  if (colstats && N != BN) {
    set_last_error("tc_gemm: fused column statistics need N in {32, 64, 128, 256} (one tile wide)");
    return GCCB_ERR_BADARG;
  }
  if (via_scratch && (!scratch || colstats)) {
    set_last_error("tc_gemm: split-K / accumulate / narrow output need a scratch buffer and cannot fuse column statistics");
    return GCCB_ERR_BADARG;
  }
  CUtensorMap ma, mb;
  int rc = make_map(&ma, A, M_cap, K, BM);
  if (rc) return rc;
  rc = make_map(&mb, B, N, K, BN);
  if (rc) return rc;
  GemmArgs g;
  g.M_cap = M_cap; g.N = N; g.K = K; g.ldo = via_scratch ? N : ldo; g.m_dev = m_dev; g.splits = splits;
  if (via_scratch) {
    g.bias = nullptr; g.alpha = 1.0f; g.out_f32 = scratch; g.out_bf16 = nullptr; g.colstats = nullptr;
  } else {
    g.bias = bias; g.alpha = alpha; g.out_f32 = out_f32; g.out_bf16 = (__nv_bfloat16*)out_bf16; g.colstats = colstats;
  }
  // the B operand stays in shared memory when it is one tile wide, at most 256 deep and reused by several row tiles
  const bool bres = !via_scratch && N == BN && K <= 256 && (BN == 256 || BN == 128) && (M_cap + BM - 1) / BM > sm_count();
  switch (BN) {
    case 256: if (bres) launch<256, true>(ma, mb, g, stream); else launch<256>(ma, mb, g, stream); break;
    case 128: if (bres) launch<128, true>(ma, mb, g, stream); else launch<128>(ma, mb, g, stream); break;
    case 64: launch<64>(ma, mb, g, stream); break;
    default: launch<32>(ma, mb, g, stream); break;
  }
  if (via_scratch) {
    const size_t total = (size_t)M_cap * n_out;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
    GCCB_LAUNCH(tc_splitk_reduce_kernel, blocks, 256, 0, stream, (const float*)scratch, splits, M_cap, n_out, N, ldo,
                m_dev, alpha, beta, bias, out_f32, (__nv_bfloat16*)out_bf16);
  }
  return check_launch("tc_gemm");
}

// fp32 -> bf16 (optionally transposed), optional per-column affine + ReLU first; zero padded to
// [rows_pad][cols_pad] (dst [cols_pad][rows_pad] when transposed); rows >= *rows_dev are zeros.
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, int rows, int cols, int lds, __nv_bfloat16* __restrict__ dst,
                 int rows_pad, int cols_pad, const int32_t* __restrict__ rows_dev,
                 const float* __restrict__ sc, const float* __restrict__ sh, int relu) {
  const int R = rows_dev ? min(max(*rows_dev, 0), rows) : rows;
  const size_t total = (size_t)rows_pad * cols_pad;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / cols_pad), c = (int)(idx - (size_t)r * cols_pad);
    float v = 0.f;
    if (r < R && c < cols) {
      v = src[(size_t)r * lds + c];
      if (sc) v = fmaf(v, sc[c], sh[c]);
      if (relu) v = fmaxf(v, 0.f);
    }
    dst[idx] = __float2bfloat16_rn(v);
  }
}

// transposing variant: one 64 x 64 tile per CTA through shared memory: 128-bit coalesced reads along the source
// rows, 32-bit (2 x bf16) coalesced writes along the destination rows; grid = all tiles (the weight-gradient
// operands of a config-4 batch are 100k+ rows: the pass has to run near HBM speed)
__global__ void __launch_bounds__(256)
cast_bf16_t_kernel(const float* __restrict__ src, int rows, int cols, int lds, __nv_bfloat16* __restrict__ dst,
                   int rows_pad, int cols_pad, const int32_t* __restrict__ rows_dev,
                   const float* __restrict__ sc, const float* __restrict__ sh, int relu) {
  __shared__ float tile[64][65];
  const int R = rows_dev ? min(max(*rows_dev, 0), rows) : rows;
  const int tcn = (cols_pad + 63) / 64;
  const int r0 = (blockIdx.x / tcn) * 64, c0 = (blockIdx.x % tcn) * 64;
  const int tid = threadIdx.x;
  const bool vec_ok = (lds % 4) == 0 && ((uintptr_t)src % 16) == 0;
  if (r0 >= R) {                                           // nothing valid in this row band: zeros
    for (int i = tid; i < 64 * 32; i += 256) {
      const int c = c0 + i / 32, r = r0 + (i % 32) * 2;
      if (c < cols_pad && r < rows_pad) *reinterpret_cast<uint32_t*>(dst + (size_t)c * rows_pad + r) = 0u;
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = tid + k * 256;                           // 1024 float4 slots: row = i / 16, 4 columns at (i % 16) * 4
    const int rr = i >> 4, cc = (i & 15) * 4;
    const int r = r0 + rr, c = c0 + cc;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < R) {
      if (vec_ok && c + 3 < cols) {
        const float4 t = *reinterpret_cast<const float4*>(src + (size_t)r * lds + c);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (c + q < cols) v[q] = src[(size_t)r * lds + c + q];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (c + q < cols) {
          if (sc) v[q] = fmaf(v[q], sc[c + q], sh[c + q]);
          if (relu) v[q] = fmaxf(v[q], 0.f);
        } else v[q] = 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) tile[rr][cc + q] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = tid + k * 256;                           // 2048 pairs: column = i / 32, rows 2 * (i % 32), +1
    const int cc = i >> 5, rr = (i & 31) * 2;
    const int c = c0 + cc, r = r0 + rr;
    if (c < cols_pad && r < rows_pad) {                    // rows_pad is even (multiple of 64)
      const __nv_bfloat162 pr = __floats2bfloat162_rn(tile[rr][cc], tile[rr + 1][cc]);
      *reinterpret_cast<uint32_t*>(dst + (size_t)c * rows_pad + r) = *reinterpret_cast<const uint32_t*>(&pr);
    }
  }
}

int cast_bf16(const float* src, int rows, int cols, int lds, void* dst, int rows_pad, int cols_pad, int transpose,
              const int32_t* rows_dev, cudaStream_t stream, const float* sc, const float* sh, int relu) {
  if (transpose && (rows_pad % 2) == 0) {
    const int tiles = ((rows_pad + 63) / 64) * ((cols_pad + 63) / 64);
    GCCB_LAUNCH(cast_bf16_t_kernel, tiles, 256, 0, stream, src, rows, cols, lds, (__nv_bfloat16*)dst, rows_pad,
                cols_pad, rows_dev, sc, sh, relu);
    return check_launch("cast_bf16 (transposed)");
  }
  if (transpose) {
    set_last_error("cast_bf16: the transposed cast needs an even rows_pad");
    return GCCB_ERR_BADARG;
  }
  const size_t total = (size_t)rows_pad * cols_pad;
  int blocks = (int)((total + 1023) / 1024);
  if (blocks > 16 * sm_count()) blocks = 16 * sm_count();
  if (blocks < 1) blocks = 1;
  GCCB_LAUNCH(cast_bf16_kernel, blocks, 256, 0, stream, src, rows, cols, lds, (__nv_bfloat16*)dst, rows_pad, cols_pad,
              rows_dev, sc, sh, relu);
  return check_launch("cast_bf16");
}

}  // namespace tc
}  // namespace gccb

using namespace gccb;

extern "C" int gccb_tc_gemm_bf16(const void* A, const void* B, int32_t M_cap, int32_t N, int32_t K,
                                 const int32_t* m_dev, const float* bias, float alpha, float* out_f32,
                                 void* out_bf16, int32_t ldo, double* colstats, int32_t splits, float* scratch,
                                 gccb_stream_t stream) {
  return tc::gemm_bf16(A, B, M_cap, N, K, m_dev, bias, alpha, out_f32, out_bf16, ldo, colstats, splits, scratch,
                       (cudaStream_t)stream);
}

extern "C" int gccb_cast_bf16(const float* src, int32_t rows, int32_t cols, int32_t lds, void* dst, int32_t rows_pad,
                              int32_t cols_pad, int32_t transpose, const int32_t* rows_dev, gccb_stream_t stream) {
  if (!src || !dst || rows < 0 || cols <= 0 || lds < cols || rows_pad < rows || cols_pad < cols) {
    set_last_error("gccb_cast_bf16: bad argument");
    return GCCB_ERR_BADARG;
  }
  return tc::cast_bf16(src, rows, cols, lds, dst, rows_pad, cols_pad, transpose, rows_dev, (cudaStream_t)stream);
}

#else   // GCCB_EMU: the CPU emulator cannot run tcgen05 / TMA; the entry points report it
extern "C" int gccb_tc_gemm_bf16(const void*, const void*, int32_t, int32_t, int32_t, const int32_t*, const float*,
                                 float, float*, void*, int32_t, double*, int32_t, float*, gccb_stream_t) {
  gccb::set_last_error("tc_gemm: tensor-core path is not available under the CPU emulator");
  return GCCB_ERR_ARCH;
}
extern "C" int gccb_cast_bf16(const float*, int32_t, int32_t, int32_t, void*, int32_t, int32_t, int32_t,
                              const int32_t*, gccb_stream_t) {
  gccb::set_last_error("cast_bf16: not available under the CPU emulator");
  return GCCB_ERR_ARCH;
}
#endif
