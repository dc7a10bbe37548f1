// Squared-L2 all-entity scoring on the 5th-generation tensor cores (tcgen05 + TMEM) with a fused
// compare-and-count epilogue: the scores never reach HBM.
//
// Replaces the score/sort/scan of common/evaluation.cpp:129-164 for the squared-L2 scorers
// (transe/transe.cpp:22-24) where the path is a dense contraction:
//     E(c) = sum_i ((V_i - C[c]_i) - d'_i)^2 = |u|^2 + n_c - 2 u.C[c],   u = V - d',  n_c = |C[c]|^2
// so "candidate c ranks before the truth" <=> s(c) = n_c - 2 u.C[c] < T = E_true - |u|^2.
// G = U x C^T is a (queries x entities x D) GEMM.  Precision: both operands are split into two
// bf16 terms (x = hi + lo, |x - hi - lo| <= 2^-17 |x|) and three products hi*hi + hi*lo + lo*hi are
// accumulated in fp32 in TMEM (relative error <= 2^-15 of sum |u_i||c_i|).  Every candidate whose
// approximate score lies within a proven band around T is re-scored in exact fp64 with the
// reference's operation order (recheck_kernel in rank.cu), so the integer ranks stay bit-exact.
//
// Kernel shape (one CTA per SM = 2 x 128 queries x a range of 128-candidate tiles; 8 epilogue warps + a copy producer +
// an MMA issuer; 229 KB of shared memory, all 512 TMEM columns):
//   A (u_hi, u_lo of both query tiles) is loaded once into shared memory, K-major, no swizzle: core matrix = 8 rows x 16 B,
//   laid out [16-byte K chunk][row] so LBO = 2048 B (next K chunk), SBO = 128 B (next 8 rows);
//   per candidate tile: B (c_hi, c_lo; |c|^2 rides in three bias columns) -> a 2-stage shared-memory ring by 1-D bulk copies;
//   one elected thread issues 21 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16) per query tile into that tile's
//   128-column TMEM accumulator (2 stages x 2 query tiles) and commits to the accumulator's own mbarrier; the four warps
//   of a query tile then read their 32 TMEM lanes (= 32 queries) with tcgen05.ld 32x32b.x32 and count / collect band
//   candidates while the tensor core works on the other accumulators.

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "rank_tc.h"

namespace kb2e {
namespace tc {

constexpr int BM = 128;          // queries per CTA tile (TMEM lanes)
constexpr int BN = 128;          // candidates per tile (TMEM columns of one accumulator)
constexpr int KC = kRowChunks;   // 16-byte chunks per operand row (8 bf16 each)
constexpr int KSTEPS = KC / 2;   // one MMA consumes K = 16 bf16 = 2 chunks
constexpr int TILE_BYTES = BM * KC * 16;   // one operand tile: [chunk][row][16 B]
constexpr int STAGES = 2;        // B-tile ring and TMEM accumulator ring
constexpr int MT = 2;            // 128-query tiles per CTA: every candidate tile fetched from L2 feeds MT accumulators
constexpr int CSPLIT = 1;        // epilogue warps per (accumulator, TMEM lane quarter): each takes BN / CSPLIT columns (2 was measured: no gain)
constexpr int EPI_WARPS = 4 * MT * CSPLIT;
constexpr int THREADS = 32 * (EPI_WARPS + 2);   // epilogue warps, + copy producer, + MMA issuer
constexpr int SMEM_BYTES = (2 * MT + 2 * STAGES) * TILE_BYTES + 128;
constexpr int TMEM_COLS = STAGES * MT * BN;      // 512: the whole tensor memory of the SM

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_NONE, version 1.
// Tile layout [16-byte K chunk][row]: core matrix = 8 rows x 16 B contiguous (128 B),
// leading byte offset (next K chunk) = 128 rows x 16 B, stride byte offset (next 8 rows) = 128 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
   uint64_t d = 0;
   d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
   d |= (uint64_t)((BM * 16) >> 4) << 16;      // leading byte offset
   d |= (uint64_t)(128 >> 4) << 32;            // stride byte offset
   d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
   return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
   asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
   uint32_t ok;
   do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
   } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
   // arrives on the mbarrier once every previously issued MMA has completed (implies fence::before_thread_sync)
   asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

// One CTA per work item = (128-query tile, range of candidate tiles).  Warp-specialised, all hand-offs by
// mbarrier: producer --full--> MMA --tmem_full--> epilogue --tmem_empty--> MMA, MMA --empty--> producer.
__global__ void __launch_bounds__(THREADS, 1) rank_l2_tc_kernel(const TcArgs a) {
   extern __shared__ __align__(128) unsigned char smem[];
   unsigned char* sA = smem;                             // MT x [hi | lo]
   unsigned char* sB = smem + 2 * MT * TILE_BYTES;       // STAGES x [hi | lo]
   uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 * MT + 2 * STAGES) * TILE_BYTES);
   uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 5 + 2 * STAGES * MT);
   const uint32_t bar_a = smem_u32(bars + 0);
   const uint32_t bar_full = smem_u32(bars + 1);         // + stage
   const uint32_t bar_empty = smem_u32(bars + 3);        // + stage
   // one pair per ACCUMULATOR (stage, query tile), not per stage: the four epilogue warps of query tile 0 start on its
   // accumulator while the tensor core is still busy with tile 1, and the MMAs of tile 0 two stages on wait only for those
   // four warps.  (With one pair per stage both sides spent ~40 % of their time waiting for each other: ncu source view.)
   const uint32_t bar_tfull = smem_u32(bars + 5);                      // + stage * MT + query tile
   const uint32_t bar_tempty = smem_u32(bars + 5 + STAGES * MT);       // + stage * MT + query tile
   static_assert((5 + 2 * STAGES * MT) * 8 + 4 <= 128, "barrier block");

   const int warp = threadIdx.x >> 5;
   const int lane = threadIdx.x & 31;
   const long long q0 = (long long)blockIdx.x * (MT * BM);
   const int tiles_total = a.n_pad / BN;
   const int t_begin = (int)((long long)tiles_total * blockIdx.y / gridDim.y);
   const int t_end = (int)((long long)tiles_total * (blockIdx.y + 1) / gridDim.y);
   const int n_iter = t_end - t_begin;
   // Every CTA walks its candidate range from a different starting tile.  With a common order all 148 SMs ask the L2 for
   // the same 57 KB at the same moment, and requests for one line are served one after the other: the operand loads then
   // take about as long as the MMAs they feed and cannot hide behind two stages (ncu source view of the common-order
   // kernel: 30 % of all stall samples are epilogue warps waiting for an accumulator, the tensor pipe is 55 % busy).
   const int rot = n_iter > 1 ? (int)(((uint32_t)blockIdx.x * 2654435761u >> 10) % (uint32_t)n_iter) : 0;
   auto tile_of = [&](int i) { const int t = i + rot; return t_begin + (t >= n_iter ? t - n_iter : t); };

   if (threadIdx.x == 0) {
      mbar_init(bar_a, 1);
      for (int s = 0; s < STAGES; s++) {
         mbar_init(bar_full + 8 * s, 1);
         mbar_init(bar_empty + 8 * s, 1);
         for (int m = 0; m < MT; m++) {
            mbar_init(bar_tfull + 8 * (s * MT + m), 1);
            mbar_init(bar_tempty + 8 * (s * MT + m), 4 * CSPLIT);
         }
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   if (warp == EPI_WARPS + 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"((uint32_t)TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
   }
   asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
   __syncthreads();
   asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
   const uint32_t tmem_base = *s_tmem;

   if (warp == EPI_WARPS) {
      // ===== copy producer: operand tiles are stored in global memory already in the shared-memory layout =====
      if (lane == 0) {
         mbar_expect_tx(bar_a, 2 * MT * TILE_BYTES);
         for (int m = 0; m < MT; m++) {
            const size_t aoff = ((size_t)blockIdx.x * MT + m) * TILE_BYTES;
            bulk_copy(smem_u32(sA + (2 * m) * TILE_BYTES), reinterpret_cast<const unsigned char*>(a.u_hi) + aoff, TILE_BYTES, bar_a);
            bulk_copy(smem_u32(sA + (2 * m + 1) * TILE_BYTES), reinterpret_cast<const unsigned char*>(a.u_lo) + aoff, TILE_BYTES, bar_a);
         }
         for (int i = 0; i < n_iter; i++) {
            const int s = i & 1;
            mbar_wait(bar_empty + 8 * s, ((i >> 1) & 1) ^ 1);   // slot free (passes at once the first time round)
            const size_t off = (size_t)tile_of(i) * TILE_BYTES;
            mbar_expect_tx(bar_full + 8 * s, 2 * TILE_BYTES);
            bulk_copy(smem_u32(sB + (2 * s) * TILE_BYTES), reinterpret_cast<const unsigned char*>(a.c_hi) + off, TILE_BYTES, bar_full + 8 * s);
            bulk_copy(smem_u32(sB + (2 * s + 1) * TILE_BYTES), reinterpret_cast<const unsigned char*>(a.c_lo) + off, TILE_BYTES, bar_full + 8 * s);
         }
      }
   } else if (warp == EPI_WARPS + 1) {
      // ===== MMA issuer: a single thread drives the tensor core =====
      if (lane == 0) {
         mbar_wait(bar_a, 0);
         for (int i = 0; i < n_iter; i++) {
            const int s = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            mbar_wait(bar_full + 8 * s, ph);         // operand bytes have landed
            const uint32_t bh = smem_u32(sB + (2 * s) * TILE_BYTES), bl = smem_u32(sB + (2 * s + 1) * TILE_BYTES);
#pragma unroll
            for (int m = 0; m < MT; m++) {
               mbar_wait(bar_tempty + 8 * (s * MT + m), ph ^ 1);   // this accumulator drained by its four epilogue warps
               asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
               const uint32_t ah = smem_u32(sA + (2 * m) * TILE_BYTES), al = smem_u32(sA + (2 * m + 1) * TILE_BYTES);
               const uint32_t acc = tmem_base + (uint32_t)((s * MT + m) * BN);
#pragma unroll
               for (int ks = 0; ks < KSTEPS; ks++) {
                  const uint32_t off = ks * 2 * (BM * 16);
                  mma_bf16(acc, make_desc(ah + off), make_desc(bh + off), ks > 0 ? 1u : 0u);   // hi * hi
                  mma_bf16(acc, make_desc(ah + off), make_desc(bl + off), 1u);                 // hi * lo
                  mma_bf16(acc, make_desc(al + off), make_desc(bh + off), 1u);                 // lo * hi
               }
               umma_commit(bar_tfull + 8 * (s * MT + m));    // this accumulator is ready for its epilogue warps
            }
            umma_commit(bar_empty + 8 * s);    // shared-memory slot can be refilled
         }
      }
   } else {
      // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) .. of accumulator (w / 4) % MT, columns [BN / CSPLIT * (w / (4 MT)) ..);
      // thread = one query.  Measured with KB2E_TC_EXPERIMENT (timing only): without the compare-and-count the kernel takes
      // 0.72 ms instead of 0.88 ms, with one TMEM load in four 0.75 ms -- the TMEM reads are free, the ~0.72 ms are the MMAs
      // (both operands from shared memory: 8 KB per 128 x 128 x 16 MMA is the shared-memory bandwidth of the SM), and the
      // compare-and-count is not fully hidden behind them; sixteen epilogue warps instead of eight changed nothing (0.883 ms).
      const int mhalf = (warp >> 2) % MT;
      const int chalf = warp / (4 * MT);
      const long long q = q0 + mhalf * BM + (warp & 3) * 32 + lane;
      const float g_lo = a.thr_lo[q];   // d >  g_lo           -> candidate certainly ranks before the truth
      const float g_hi = a.thr_hi[q];   // g_hi <= d <= g_lo   -> undecided: exact fp64 recheck
      // The hot loop sees every accumulator value once, so it is kept to 2.5 independent instructions per value (compare +
      // predicated add, one three-input min of |.| per two values; the straightforward
      // `less += d > g_lo; band |= d >= g_hi && !(d > g_lo)` compiled to 7 with a serial chain and made the kernel
      // epilogue-bound: 1.22 -> 0.97 ms):  "certainly before" <=> d > T;  "a band candidate may be in this chunk" <=>
      // min |d| <= T, with T = max(|g_lo|, |g_hi|), so that d > T implies d > g_lo and d in [g_hi, g_lo] implies |d| <= T.
      // Values between g_lo and T are merely sent to the exact recheck as well.  (Loading the whole 128-column row with two
      // x64 TMEM loads before one wait was measured 2.3x SLOWER than four x32 load / wait / process rounds.)
      // (g_lo, g_hi) straddle zero almost symmetrically: the query's threshold is part of the accumulator (prep_queries_kernel)
      const float T = fmaxf(fabsf(g_lo), fabsf(g_hi)) + 1.0e-37f;
      int less = 0;
      for (int i = 0; i < n_iter; i++) {
         const int s = i & 1;
         mbar_wait(bar_tfull + 8 * (s * MT + mhalf), (i >> 1) & 1);
         asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
         const long long c0 = (long long)tile_of(i) * BN;
#pragma unroll 1
         for (int cb = chalf * (BN / CSPLIT); cb < (chalf + 1) * (BN / CSPLIT); cb += 32) {
            uint32_t v[32];
            if (a.experiment == 2 && (cb & 96) != 0) continue;
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((s * MT + mhalf) * BN + cb);
            asm volatile(
               "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (a.experiment == 1) { less += (int)(v[0] & 1u); continue; }
            int l0 = 0, l1 = 0, l2 = 0, l3 = 0;
            float m0 = 3.0e38f, m1 = 3.0e38f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
               const float d0 = __uint_as_float(v[j]), d1 = __uint_as_float(v[j + 1]);
               const float d2 = __uint_as_float(v[j + 2]), d3 = __uint_as_float(v[j + 3]);
               asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(l0) : "f"(d0), "f"(T));
               asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(l1) : "f"(d1), "f"(T));
               asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(l2) : "f"(d2), "f"(T));
               asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(l3) : "f"(d3), "f"(T));
               m0 = fminf(m0, fminf(fabsf(d0), fabsf(d1)));
               m1 = fminf(m1, fminf(fabsf(d2), fabsf(d3)));
            }
            less += (l0 + l1) + (l2 + l3);
            if (fminf(m0, m1) <= T) {   // rare: collect the undecided candidates of this 32-column chunk
#pragma unroll
               for (int j = 0; j < 32; j++) {
                  const float d = __uint_as_float(v[j]);
                  if (fabsf(d) <= T) {
                     unsigned int slot = atomicAdd(a.band_count, 1u);
                     if (slot < a.band_cap) a.band[slot] = make_int2((int)q, (int)(c0 + cb + j));
                  }
               }
            }
         }
         // this warp has finished reading the accumulator
         asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
         __syncwarp();
         if (lane == 0) mbar_arrive(bar_tempty + 8 * (s * MT + mhalf));
      }
      if (q < a.nq) {
         if (less) atomicAdd(a.q_less + q, less);   // (CSPLIT warps and gridDim.y CTAs share a query; q_less is zeroed by the host)
      }
   }
   asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
   __syncthreads();
   if (warp == EPI_WARPS + 1) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
   }
}

// ---- operand preparation ---------------------------------------------------------------------------
// Operand tiles live in global memory in the exact shared-memory layout: tile of 128 rows =
// [16-byte K chunk (14)][row (128)][8 bf16], so a tile is one contiguous 28,672-byte bulk copy.
__device__ __forceinline__ size_t tiled_index(long long row, int i) {
   const long long tile = row / BM;
   const int r = (int)(row - tile * BM);
   return (size_t)tile * (TILE_BYTES / 2) + (size_t)(i >> 3) * (BM * 8) + (size_t)r * 8 + (i & 7);
}

__device__ __forceinline__ void split_bf16(double x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
   float xf = (float)x;
   hi = __float2bfloat16_rn(xf);
   lo = __float2bfloat16_rn(xf - __bfloat162float(hi));
}

// candidates: C (fp64 [n][D] row-major) -> tiled c_hi / c_lo; columns D..D+2 of c_hi carry -|c|^2/2 as three
// bf16 terms (the query side holds 1.0 there), so the accumulator is g = u.c - |c|^2/2 and
// "score < T" becomes g > -T/2 with no per-column work in the epilogue.  Also max |c| for the band.
__global__ void prep_candidates_kernel(const double* __restrict__ c64, int n, int n_pad, int D,
                                       __nv_bfloat16* c_hi, __nv_bfloat16* c_lo, unsigned int* cmax_bits) {
   int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   int lane = threadIdx.x & 31;
   if (row >= n_pad) return;
   double s = 0.0;
   for (int i = lane; i < KC * 8; i += 32) {
      double x = (row < n && i < D) ? c64[(size_t)row * D + i] : 0.0;
      __nv_bfloat16 h, l;
      split_bf16(x, h, l);
      c_hi[tiled_index(row, i)] = h;
      c_lo[tiled_index(row, i)] = l;
      s += x * x;
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
   __syncwarp();   // the loop's zeros in the bias columns are overwritten below by lane 0
   if (lane == 0) {
      // padding rows get a huge negative bias: they never rank before anything (finite, so 0 * bias stays 0)
      float b = row < n ? (float)(-0.5 * s) : -1.0e30f;
      __nv_bfloat16 b0 = __float2bfloat16_rn(b);
      float r1 = b - __bfloat162float(b0);
      __nv_bfloat16 b1 = __float2bfloat16_rn(r1);
      __nv_bfloat16 b2 = __float2bfloat16_rn(r1 - __bfloat162float(b1));
      c_hi[tiled_index(row, D + 0)] = b0;
      c_hi[tiled_index(row, D + 1)] = b1;
      c_hi[tiled_index(row, D + 2)] = b2;
      // columns D+3 .. D+5: 1.0 against the query's threshold terms (prep_queries_kernel)
      const __nv_bfloat16 one = __float2bfloat16_rn(1.0f);
      c_hi[tiled_index(row, D + 3)] = one;
      c_hi[tiled_index(row, D + 4)] = one;
      c_hi[tiled_index(row, D + 5)] = one;
      if (row < n) atomicMax(cmax_bits, __float_as_uint((float)sqrt(s) * 1.0000002f));
   }
}

// queries: u = V - d' (exact fp64) -> tiled u_hi / u_lo (+ 1.0 in the three bias columns);
// thresholds on g from E_true and |u|^2:  E(c) < E_true  <=>  g = u.c - |c|^2/2 > (|u|^2 - E_true) / 2
__global__ void prep_queries_kernel(const double* __restrict__ c64, const double* __restrict__ rel64,
                                    const int32_t* q_fixed, const int32_t* q_rel, const int32_t* q_side, const double* q_etrue,
                                    long long nq, long long q_pad, int D, const unsigned int* cmax_bits,
                                    __nv_bfloat16* u_hi, __nv_bfloat16* u_lo, float* thr_lo, float* thr_hi) {
   long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   int lane = threadIdx.x & 31;
   if (q >= q_pad) return;
   double s = 0.0;
   const bool real = q < nq;
   const double* v = real ? c64 + (size_t)q_fixed[q] * D : nullptr;
   const double* d = real ? rel64 + (size_t)q_rel[q] * D : nullptr;
   const double sign = real && q_side[q] ? -1.0 : 1.0;
   for (int i = lane; i < KC * 8; i += 32) {
      double x = (real && i < D) ? v[i] - sign * d[i] : 0.0;
      __nv_bfloat16 h, l;
      split_bf16(x, h, l);
      if (real && i >= D && i < D + 3) h = __float2bfloat16_rn(1.0f);
      u_hi[tiled_index(q, i)] = h;
      u_lo[tiled_index(q, i)] = l;
      s += x * x;
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
   __syncwarp();   // the loop's zeros in the threshold columns are overwritten below by lane 0
   if (lane == 0) {
      if (!real) {
         thr_lo[q] = __int_as_float(0x7fc00000);  // NaN: every comparison of a padding row is false -- it never counts and
         thr_hi[q] = __int_as_float(0x7fc00000);  // never sends a candidate to the recheck
      } else {
         const double cmax = (double)__uint_as_float(*cmax_bits);
         const double G = 0.5 * (s - q_etrue[q]);
         // The threshold itself goes INTO the GEMM: columns D+3 .. D+5 of the query carry -G as three bf16 terms (the
         // candidates hold 1.0 there), so the accumulator is d = u.c - |c|^2/2 - G~ and the epilogue tests d against +-T
         // with no per-value subtraction.  G~ = the exact sum of the three terms; T absorbs G - G~.
         const float gneg = (float)(-G);
         const __nv_bfloat16 t0 = __float2bfloat16_rn(gneg);
         const float r1 = gneg - __bfloat162float(t0);
         const __nv_bfloat16 t1 = __float2bfloat16_rn(r1);
         const __nv_bfloat16 t2 = __float2bfloat16_rn(r1 - __bfloat162float(t1));
         u_hi[tiled_index(q, D + 3)] = t0;
         u_hi[tiled_index(q, D + 4)] = t1;
         u_hi[tiled_index(q, D + 5)] = t2;
         const double Gt = -((double)__bfloat162float(t0) + (double)__bfloat162float(t1) + (double)__bfloat162float(t2));
         // |d~ - d| <= 1.9e-5 |u||c| (three-product bf16 split) + ~2e-5 (|u||c| + |G|) (fp32 accumulation of 345
         // products) + 3 bf16 terms of |c|^2/2 (<= 2^-25 |c|^2); |G| <= (|u| + |c|max)^2 / 2.  The band
         // 2^-13 |u||c|max + eps is > 2x that.
         const double unorm = sqrt(s);
         const double delta = 4.0 * (1.0 / 32768.0) * unorm * cmax + 3e-6 * (1.0 + unorm * cmax + cmax * cmax) +
                              2e-6 * (unorm + cmax) * (unorm + cmax);
         // certainly before the truth: g~ > G + delta  <=>  d > (G - G~) + delta;  undecided: |d - (G - G~)| <= delta
         thr_lo[q] = __double2float_ru((G - Gt) + delta);
         thr_hi[q] = __double2float_rd((G - Gt) - delta);
      }
   }
}

// Filter pass, first cut: the (query, known-true neighbour) pairs are scored from the SAME bf16 hi / lo operand tiles and
// against the same thresholds as the tensor-core kernel -- d = sum_k (u_hi + u_lo)_k (c_hi + c_lo)_k over all columns, the
// bias columns included, in fp32 (8-column dots added chunk by chunk: error (8 + 14) 2^-24 of sum |terms|, well inside the
// band the thresholds carry; the lo * lo products the MMA leaves out only bring d closer to the exact value).
// d > T: the neighbour certainly ranks before the truth -> counted here; d < -T: certainly not; both are struck from the
// list (candidate = -1), so that the exact fp64 filter_pairs_kernel that follows scores only the undecided few.
// 896 bytes of operands per pair instead of 2,400, no chain of dependent fp64 additions.
__global__ void filter_prefilter_kernel(const __nv_bfloat16* __restrict__ u_hi, const __nv_bfloat16* __restrict__ u_lo,
                                        const __nv_bfloat16* __restrict__ c_hi, const __nv_bfloat16* __restrict__ c_lo,
                                        const float* __restrict__ thr_lo, const float* __restrict__ thr_hi, long long q_base,
                                        int2* pairs, const unsigned int* __restrict__ pair_count, unsigned int pair_cap,
                                        int32_t* q_filt_less) {
   const unsigned int n = min(*pair_count, pair_cap);
   for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
      const int2 pr = pairs[k];
      if (pr.y < 0) continue;
      const long long q = (long long)pr.x - q_base;
      const float T = fmaxf(fabsf(thr_lo[q]), fabsf(thr_hi[q]));
      float d = 0.f;
#pragma unroll 2
      for (int ch = 0; ch < KC; ch++) {
         const uint4 uh = *reinterpret_cast<const uint4*>(u_hi + tiled_index(q, 8 * ch));
         const uint4 ul = *reinterpret_cast<const uint4*>(u_lo + tiled_index(q, 8 * ch));
         const uint4 xh = *reinterpret_cast<const uint4*>(c_hi + tiled_index(pr.y, 8 * ch));
         const uint4 xl = *reinterpret_cast<const uint4*>(c_lo + tiled_index(pr.y, 8 * ch));
         const uint32_t a4[4] = {uh.x, uh.y, uh.z, uh.w}, b4[4] = {ul.x, ul.y, ul.z, ul.w};
         const uint32_t c4[4] = {xh.x, xh.y, xh.z, xh.w}, d4[4] = {xl.x, xl.y, xl.z, xl.w};
         float part = 0.f;
#pragma unroll
         for (int j = 0; j < 4; j++) {
            // a bf16 is the upper half of an fp32: element 2j in the low 16 bits, 2j + 1 in the high 16 bits
            const float u0 = __uint_as_float(a4[j] << 16) + __uint_as_float(b4[j] << 16);
            const float u1 = __uint_as_float(a4[j] & 0xffff0000u) + __uint_as_float(b4[j] & 0xffff0000u);
            const float x0 = __uint_as_float(c4[j] << 16) + __uint_as_float(d4[j] << 16);
            const float x1 = __uint_as_float(c4[j] & 0xffff0000u) + __uint_as_float(d4[j] & 0xffff0000u);
            part = fmaf(u0, x0, part);
            part = fmaf(u1, x1, part);
         }
         d += part;
      }
      if (d > T) {
         atomicAdd(q_filt_less + pr.x, 1);
         pairs[k].y = -1;
      } else if (d < -T) {
         pairs[k].y = -1;
      }
   }
}

}  // namespace tc

// ---- host -------------------------------------------------------------------------------------------
int tc_filter_prefilter(kb2e_ctx* c, TcState* s, long long q_base, int2* pairs, const unsigned int* pair_count, unsigned int pair_cap,
                        int32_t* q_filt_less, cudaStream_t stream) {
   tc::filter_prefilter_kernel<<<8 * c->num_sms, 256, 0, stream>>>(
      (const __nv_bfloat16*)s->u_hi, (const __nv_bfloat16*)s->u_lo, (const __nv_bfloat16*)s->c_hi, (const __nv_bfloat16*)s->c_lo,
      s->thr_lo, s->thr_hi, q_base, pairs, pair_count, pair_cap, q_filt_less);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

static inline unsigned nblk2(long long n, int t) { return (unsigned)((n + t - 1) / t); }

bool tc_supported(const kb2e_ctx* c) {
   return c->cfg.model == KB2E_MODEL_TRANSE && c->cfg.distance == KB2E_DISTANCE_L2 && c->D + 6 <= tc::kRowChunks * 8 &&
          !(c->cfg.flags & KB2E_FLAG_RANK_EXACT_ONLY);
}

int tc_prepare_candidates(kb2e_ctx* c, TcState* s) {
   const int n_pad = ((c->nE + tc::BN - 1) / tc::BN) * tc::BN;
   if (n_pad != s->n_pad) {
      pool_free(c, s->c_hi); pool_free(c, s->c_lo);
      s->c_hi = s->c_lo = nullptr;
      KB2E_CUDA(c, pool_alloc(c, &s->c_hi, (size_t)n_pad * tc::kRowChunks * 16));
      KB2E_CUDA(c, pool_alloc(c, &s->c_lo, (size_t)n_pad * tc::kRowChunks * 16));
      s->n_pad = n_pad;
   }
   if (!s->scalars) KB2E_CUDA(c, pool_alloc(c, &s->scalars, 4 * sizeof(unsigned int)));
   KB2E_CUDA(c, cudaMemsetAsync(s->scalars, 0, 4 * sizeof(unsigned int), c->stream));
   tc::prep_candidates_kernel<<<nblk2((long long)n_pad * 32, 256), 256, 0, c->stream>>>(
      c->ent64, c->nE, n_pad, c->D, (__nv_bfloat16*)s->c_hi, (__nv_bfloat16*)s->c_lo, s->scalars);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int tc_run(kb2e_ctx* c, TcState* s, const int32_t* q_fixed, const int32_t* q_rel, const int32_t* q_side, const double* q_etrue,
           long long nq, int32_t* q_less) {
   const long long q_pad = ((nq + tc::MT * tc::BM - 1) / (tc::MT * tc::BM)) * (tc::MT * tc::BM);
   if (q_pad > s->q_cap) {
      pool_free(c, s->u_hi); pool_free(c, s->u_lo); pool_free(c, s->thr_lo); pool_free(c, s->thr_hi); pool_free(c, s->band);
      s->u_hi = s->u_lo = nullptr; s->thr_lo = s->thr_hi = nullptr; s->band = nullptr;
      s->q_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &s->u_hi, (size_t)q_pad * tc::kRowChunks * 16));
      KB2E_CUDA(c, pool_alloc(c, &s->u_lo, (size_t)q_pad * tc::kRowChunks * 16));
      KB2E_CUDA(c, pool_alloc(c, &s->thr_lo, (size_t)q_pad * sizeof(float)));
      KB2E_CUDA(c, pool_alloc(c, &s->thr_hi, (size_t)q_pad * sizeof(float)));
      s->band_cap = (unsigned int)std::min<long long>(std::max<long long>(1 << 20, 64 * q_pad), 1ll << 28);
      KB2E_CUDA(c, pool_alloc(c, &s->band, (size_t)s->band_cap * sizeof(int2)));
      s->q_cap = q_pad;
   }
   unsigned int* band_count = s->scalars + 1;
   KB2E_CUDA(c, cudaMemsetAsync(band_count, 0, sizeof(unsigned int), c->stream));
   tc::prep_queries_kernel<<<nblk2(q_pad * 32, 256), 256, 0, c->stream>>>(
      c->ent64, c->rel64, q_fixed, q_rel, q_side, q_etrue, nq, q_pad, c->D, s->scalars,
      (__nv_bfloat16*)s->u_hi, (__nv_bfloat16*)s->u_lo, s->thr_lo, s->thr_hi);
   TcArgs a;
   a.u_hi = (const __nv_bfloat16*)s->u_hi; a.u_lo = (const __nv_bfloat16*)s->u_lo;
   a.c_hi = (const __nv_bfloat16*)s->c_hi; a.c_lo = (const __nv_bfloat16*)s->c_lo;
   a.thr_lo = s->thr_lo; a.thr_hi = s->thr_hi;
   a.q_less = q_less; a.band = s->band; a.band_count = band_count; a.band_cap = s->band_cap;
   a.nq = nq; a.n_pad = s->n_pad;
   { const char* e = getenv("KB2E_TC_EXPERIMENT"); a.experiment = e ? atoi(e) : 0; }
   const unsigned mtiles = (unsigned)(q_pad / (tc::MT * tc::BM));
   const int ntiles = s->n_pad / tc::BN;
   // one CTA per SM: cut the candidate range so that the work items fill whole waves (tail < 5 %)
   long long splits = 1;
   while (splits < ntiles) {
      const long long items = (long long)mtiles * splits;
      const long long waves = (items + c->num_sms - 1) / c->num_sms;
      if (items >= 2 * c->num_sms && (double)items / (waves * c->num_sms) > 0.95) break;
      splits++;
   }
   splits = std::min<long long>(splits, ntiles);
   KB2E_CUDA(c, cudaFuncSetAttribute(tc::rank_l2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
   KB2E_CUDA(c, cudaMemsetAsync(q_less, 0, (size_t)nq * sizeof(int32_t), c->stream));
   KB2E_CUDA(c, cudaEventRecord(s->e0, c->stream));
   tc::rank_l2_tc_kernel<<<dim3(mtiles, (unsigned)splits), tc::THREADS, tc::SMEM_BYTES, c->stream>>>(a);
   KB2E_CUDA(c, cudaEventRecord(s->e1, c->stream));
   KB2E_CUDA(c, cudaGetLastError());
   KB2E_CUDA(c, cudaMemcpyAsync(s->host_count, band_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
   return KB2E_OK;
}

int tc_collect(kb2e_ctx* c, TcState* s, bool* overflow) {
   float ms = 0.f;
   KB2E_CUDA(c, cudaEventElapsedTime(&ms, s->e0, s->e1));
   s->last_ms = ms;
   s->last_band = *s->host_count;
   *overflow = s->last_band > s->band_cap;
   return KB2E_OK;
}

int tc_init(kb2e_ctx* c, TcState* s) {
   if (!s->e0) {
      KB2E_CUDA(c, cudaEventCreate(&s->e0));
      KB2E_CUDA(c, cudaEventCreate(&s->e1));
      KB2E_CUDA(c, cudaMallocHost(&s->host_count, sizeof(unsigned int)));
   }
   return KB2E_OK;
}

void tc_free(kb2e_ctx* c, TcState* s) {
   pool_free(c, s->c_hi); pool_free(c, s->c_lo); pool_free(c, s->u_hi); pool_free(c, s->u_lo);
   pool_free(c, s->thr_lo); pool_free(c, s->thr_hi); pool_free(c, s->band); pool_free(c, s->scalars);
   if (s->e0) { cudaEventDestroy(s->e0); cudaEventDestroy(s->e1); cudaFreeHost(s->host_count); }
   *s = TcState();
}

}  // namespace kb2e
