// NNEDI3 on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// What the reference does per interpolated pixel (nnedi3-nns16-win8x4.hook:20-50): window mean / stddev,
// 2*nns dot products of length K = 8*S against constant weights, softmax(exp) . elliott epilogue.  That is a
// dense contraction  D[pixels x 2*nns] = A[pixels x K] . B[K x 2*nns]  with a per-row epilogue:
//
//   A  built by CUDA cores: (x - mean) * inv_std -> binary16, written in the UMMA canonical K-major
//      (SWIZZLE_NONE) layout [K/8][128 rows][8 halfs]; mean removal BEFORE the fp16 rounding is what keeps
//      the result inside the 1e-3 bound;
//   B  [K/8][2*nns][8 halfs] binary16, resident in shared memory for the whole kernel, fetched once per CTA
//      with one TMA bulk copy (cp.async.bulk -> mbarrier); rows interleave (W1_n * log2e, W2_n);
//      the biases ride along as one extra K step (A columns 1,1,1,0..; B rows bias hi/mid/lo);
//   D  fp32 in tensor memory: 128 lanes (pixels) x 128-column chunks (64 neurons); the MMA of chunk c+1 is
//      issued as soon as the last columns of chunk c have been loaded into registers;
//   epilogue: tcgen05.ld 32x32b (thread == pixel == TMEM lane, next load in flight while the current
//      one is consumed) -> ex2 (MUFU), reciprocal (MUFU or two Newton steps on the FMA pipe), running sums.
//
// Four independent warpgroups per CTA (128 threads each) walk 32x4-pixel tiles; each owns a quarter of
// TMEM (128 columns) and issues its own tcgen05.mma from one elected thread; the source window of the
// next tile is staged with cp.async while the current tile's epilogue runs.
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "tma.cuh"

namespace mpvp {
namespace {

struct NnTcArgs {
  const void* __restrict__ in;   // planes of format io.in_fmt
  void* __restrict__ out;        // planes of format io.out_fmt
  IoFmt io;
  const void* __restrict__ b_packed;  // [K/8][N][8] binary16
  const float* __restrict__ bias;     // [N] interleaved (b1*log2e, b2)
  int n, h, w;
  int group;  // neurons per accumulator block in b_packed (mpvp_weights::nn_group)
  int64_t in_sn, in_sy, out_sn, out_sy;
  int tiles_x, tiles_y;
  long long total_tiles;
};

constexpr int kTileW = 32, kTileH = 4;  // 128 pixels = 128 TMEM lanes
#ifndef MPVP_X_NN_WG
#define MPVP_X_NN_WG 4
#endif
#ifndef MPVP_X_NN_PPW
#define MPVP_X_NN_PPW 64   // ping-pong slot width (experiment: 128 with MPVP_X_NN_WG=2, i.e. 256 TMEM columns per warpgroup)
#endif
constexpr int kWG = MPVP_X_NN_WG;       // warpgroups per CTA
constexpr int kWGCols = 512 / kWG;      // TMEM columns per warpgroup
static_assert(2 * MPVP_X_NN_PPW <= kWGCols, "two ping-pong slots per warpgroup");
constexpr int kThreads = 128 * kWG;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wg_barrier(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
// issue only (no wait): the caller overlaps the load with arithmetic and calls tmem_ld_wait() before use
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B contiguous;
// SBO = byte distance between 8-row groups, LBO = byte distance between 8-element K chunks.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for sm_100
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// kind::f16 instruction descriptor: D fp32, A/B fp16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 1/u for u >= 1 on the FMA/ALU pipes: bit-trick seed (|rel err| < 0.13) + two Newton steps (< 3e-4, always
// from below).  Keeps the MUFU pipe, the binding resource of the epilogue, for the exponentials.
__device__ __forceinline__ float rcp_newton(float u) {
  float r = __int_as_float(0x7EF311C7 - __float_as_int(u));
  r = r * fmaf(-u, r, 2.0f);
  r = r * fmaf(-u, r, 2.0f);
  return r;
}

// TMA: the source window of a tile (32 x 4 pixels + halo) arrives by cp.async.bulk.tensor into the warpgroup's staging
// buffer (float32 planes that meet the 16-byte rules, tma.cuh): one elected thread per warpgroup issues the box of the
// NEXT tile while the current tile's epilogue runs, completion on a per-warpgroup mbarrier; the box starts 4 texels left
// of the tile (16-byte aligned origin) and is 40 texels wide; border tiles are patched to clamp-to-edge after arrival.
template <int S, int DIR, int NNS, int EPI, bool TMA>
__global__ void __launch_bounds__(kThreads, 1) nnedi3_tc_kernel(const __grid_constant__ NnTcArgs A, const __grid_constant__ CUtensorMap tmap) {
  constexpr int K = 8 * S;                  // window samples
  constexpr int KX = K + 16;                // + one MMA K-step carrying the biases
  constexpr int KC = K / 8;                 // 16-byte K chunks written per tile
  constexpr int N = 2 * NNS;                // accumulator columns per pixel
  // PP (ping-pong, networks with >= 128 accumulator columns): the warpgroup's 128 TMEM columns are two 64-column slots; a
  // chunk's 64 columns are loaded into registers at the top of the chunk, which frees its slot for the MMA after next, and
  // the MMA of the NEXT chunk is issued at the same point into the other slot -- so no warp ever waits in a warpgroup
  // barrier inside a tile, and the warps of a warpgroup (which sit on four different SM sub-partitions) may drift by up to
  // a chunk.  Stall sampling of the barrier-synchronised schedule: 17 % of all warp samples sit in the chunk-boundary
  // barrier, the wait for the MMA it releases, and the tile-start wait (profiles/r02 notes in DESIGN.md 4.5).
  // MEASURED SLOWER and compiled out: nns256-win8x6 6.01 vs 4.67 ms, nns128-win8x4 3.01 vs 2.51, nns64-win8x6 2.01 vs 1.75.
  // An SS-mode tcgen05.mma re-reads its A operand from shared memory for every instruction: per K = 16 step 4 KB of A
  // + N * 32 B of B.  At N = 64 that is 6 KB per 32 tensor cycles = 192 B/clk against the 128 B/clk of an SM's shared
  // memory; at N = 128 (the barrier-synchronised schedule) 8 KB per 64 cycles = exactly 128 B/clk.  Small-N MMAs are
  // shared-memory bound and starve the LDS / STS of the other warpgroups; a ping-pong schedule needs the A operand in
  // tensor memory (tcgen05.mma with A from TMEM), for which 4 x (128 accumulator + 32 A) columns do not fit in 512.
#ifndef MPVP_X_NN_PP
#define MPVP_X_NN_PP 0
#endif
  constexpr bool PP = MPVP_X_NN_PP && N >= 2 * MPVP_X_NN_PPW && EPI == 3;
  constexpr int CN = PP ? MPVP_X_NN_PPW : (N < 128 ? N : 128);     // columns per MMA chunk
  constexpr int NCH = N / CN;
  // Small networks (nns16 / nns32: one chunk of <= 64 columns per tile) keep TWO tiles in flight per warpgroup: the A
  // operand and the MMA of tile t+1 are issued before the epilogue of tile t starts (two A buffers, two TMEM slots,
  // two mbarriers), so the build -> MMA -> epilogue latency chain of a tile, which dominates when the epilogue is
  // only 16-32 neurons long, overlaps the neighbouring tile's work.
  constexpr bool DB = (NCH == 1 && CN <= 64);
  // Larger networks pipeline ACROSS tiles instead (XT): the A operand of tile t+1 is built (second A buffer) while the MMA of
  // tile t's last chunk runs, and its first MMA is issued at the barrier that ends tile t's last TMEM read -- the
  // stage-wait -> build -> barrier -> MMA -> wait chain of a tile's start is off the warpgroup's critical path.
#ifndef MPVP_X_NN_XT
#define MPVP_X_NN_XT 0   // measured slower: nns256-win8x6 4.89 vs 4.69 ms, nns128-win8x4 2.64 vs 2.51 ms (kept as an A/B knob)
#endif
  constexpr bool XT = !DB && MPVP_X_NN_XT;
  constexpr int NA = (DB || XT) ? 2 : 1;    // A buffers per warpgroup
  constexpr int HX = DIR == 0 ? 8 : S, HY = DIR == 0 ? S : 8;
  constexpr int OX = DIR == 0 ? 3 : (S / 2 - 1), OY = DIR == 0 ? (S / 2 - 1) : 3;
  constexpr int XO = TMA ? 4 : OX;                         // staged columns left of the tile
  constexpr int SW = TMA ? 40 : (kTileW + HX - 1), SH = kTileH + HY - 1;
  static_assert(!TMA || (4 + kTileW + HX - 1 - OX) <= 40, "TMA box too narrow");
  constexpr int STG = (SW * SH + 31) & ~31;                // floats per staging buffer (128-byte multiple)
  constexpr uint32_t kBBytes = (uint32_t)KX * N * 2;
  constexpr uint32_t kABytes = (uint32_t)KX * 128 * 2;

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_b = smem;                               // B operand
  unsigned char* s_a = s_b + kBBytes;                      // A operand, one per warpgroup
  float* s_stage = reinterpret_cast<float*>(s_a + kWG * NA * kABytes);  // [kWG][STG]
  uint64_t* s_mbar = reinterpret_cast<uint64_t*>(s_stage + kWG * STG);  // [kWG][2] MMA done, + 1 (B operand), + [kWG] staging, + [kWG][2] slot free (PP)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_mbar + 5 * kWG + 1);

  const int tid = threadIdx.x;
  const int wg = tid >> 7;       // warpgroup
  const int lt = tid & 127;      // pixel / TMEM lane within the tile
  const int warp = tid >> 5;
  const int tx = lt & 31, ty = lt >> 5;

  const uint32_t mbar_b = smem_u32(s_mbar + 2 * kWG);
  if (tid == 0) {
    for (int i = 0; i < 2 * kWG; ++i) mbar_init(smem_u32(s_mbar + i), 1);
    mbar_init(mbar_b, 1);
    for (int i = 0; i < kWG; ++i) mbar_init(smem_u32(s_mbar + 2 * kWG + 1 + i), 1);
    for (int i = 0; i < 2 * kWG; ++i) mbar_init(smem_u32(s_mbar + 3 * kWG + 1 + i), 4);   // one arrival per warp
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 512);
  unsigned char* my_a = s_a + wg * NA * kABytes;
  {
    // constant tail of every A row: the bias step (1, 1, 1, 0, 0, 0, 0, 0 | 0 x 8)
    const __half one = __float2half_rn(1.0f), zero = __float2half_rn(0.0f);
    __half2 h0 = __halves2half2(one, one), h1 = __halves2half2(one, zero), hz = __halves2half2(zero, zero);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&hz); pk.w = pk.z;
#pragma unroll
    for (int bufi = 0; bufi < NA; ++bufi) *reinterpret_cast<uint4*>(my_a + bufi * kABytes + KC * (128 * 16) + lt * 16) = pk;
    pk.x = pk.y = pk.z;
#pragma unroll
    for (int bufi = 0; bufi < NA; ++bufi) *reinterpret_cast<uint4*>(my_a + bufi * kABytes + (KC + 1) * (128 * 16) + lt * 16) = pk;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_expect_tx(mbar_b, kBBytes);
    bulk_g2s(smem_u32(s_b), A.b_packed, kBBytes, mbar_b);
  }
  const uint32_t tmem_base = *s_tmem;
  mbar_wait(mbar_b, 0);

  float* my_stage = s_stage + wg * STG;
  const uint32_t a_addr = smem_u32(my_a), b_addr = smem_u32(s_b);
  const uint32_t idesc = make_idesc(CN);
  const uint32_t my_mbar = smem_u32(s_mbar + 2 * wg);
  const uint32_t d_col = tmem_base + (uint32_t)(wg * kWGCols);
  const uint32_t d_lane = d_col + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0;  // parity of this warpgroup's MMA-done barrier
  const uint32_t free_bar = smem_u32(s_mbar + 3 * kWG + 1 + 2 * wg);   // PP: [2] "slot loaded into registers by all four warps"
  uint32_t full_par = 0, free_par = 0;                                  // PP: parity bits of the two full / free barriers
  const uint32_t stage_bar = smem_u32(s_mbar + 2 * kWG + 1 + wg);
  uint32_t stage_phase = 0;   // parity of this warpgroup's staging barrier (TMA)
  const uint64_t tmap_ptr = reinterpret_cast<uint64_t>(&tmap);

  auto issue_chunk = [&](int c, int slot = 0, int aslot = -1) {
    // D[128 x CN] = A[128 x KX] . B[rows c*CN .. c*CN+CN-1]^T   (slot: TMEM slot / mbarrier of the DB mode; aslot: A buffer,
    // = slot unless given)
    if (aslot < 0) aslot = slot;
#pragma unroll
    for (int j = 0; j < KX / 16; ++j) {
      const uint64_t ad = make_desc(a_addr + aslot * kABytes + j * 2 * (128 * 16), 128 * 16, 128);
      const uint64_t bd = make_desc(b_addr + c * CN * 16 + j * 2 * (N * 16), N * 16, 128);
      umma_f16(d_col + (uint32_t)(slot * CN), ad, bd, idesc, j > 0 ? 1u : 0u);
    }
    umma_commit(my_mbar + 8u * (uint32_t)slot);
  };

  // asynchronous staging of a tile's source rectangle (clamp-to-edge): cp.async, no registers held, so the
  // loads of tile t+1 fly under the epilogue of tile t
  auto prefetch_tile = [&](long long tile, const TileWalk& tw) {
    if (tile >= A.total_tiles) return;
    const int f = tw.f;
    const int x0 = tw.tix * kTileW, y0 = tw.tiy * kTileH;
    if constexpr (TMA) {
      if (lt == 0) {
        fence_proxy_async();   // the warpgroup's reads of the buffer (ordered by its barrier) precede the async write
        tma_expect(stage_bar, SW * SH * 4);
        tma_load_3d(smem_u32(my_stage), tmap_ptr, x0 - XO, y0 - OY, f, stage_bar);
      }
      return;
    }
    const int64_t src0 = (int64_t)f * A.in_sn;
    for (int i = lt; i < SW * SH; i += 128) {
      const int sy = i / SW, sx = i - sy * SW;
      const int gx = clampi(x0 + sx - OX, 0, A.w - 1), gy = clampi(y0 + sy - OY, 0, A.h - 1);
      const int64_t off = src0 + (int64_t)gy * A.in_sy + gx;
      if (A.io.in_fmt == MPVP_FMT_F32) {
        const uint32_t dst = smem_u32(my_stage + i);
        const float* g = static_cast<const float*>(A.in) + off;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(g) : "memory");
      } else {  // other plane formats are converted on the way in (synchronous loads)
        my_stage[i] = load_px(A.in, off, A.io.in_fmt, A.io.in_max);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // im2col + normalisation of the staged window -> A operand in `abuf`; returns the pixel's mean, stddev and centre sample
  auto build_a = [&](unsigned char* abuf, float& mstd0, float& mstd1, float& orig) {
    // packed f32x2 throughout: two window samples per FADD2 / FFMA2
    float2 xs[K / 2];
    float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; k += 2) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int a = (k + e) / S, b = (k + e) % S;
        const int dx = DIR == 0 ? a : b, dy = DIR == 0 ? b : a;
        v[e] = my_stage[(ty + dy) * SW + tx + dx + (XO - OX)];
      }
      xs[k / 2] = make_float2(v[0], v[1]);
      sum2 = __fadd2_rn(sum2, xs[k / 2]);
      sq2 = __ffma2_rn(xs[k / 2], xs[k / 2], sq2);
    }
    mstd0 = (sum2.x + sum2.y) / (float)K;
    mstd1 = (sq2.x + sq2.y) / (float)K - mstd0 * mstd0;
    const float mstd2 = mstd1 >= kEps ? rsqrtf(mstd1) : 0.0f;
    mstd1 *= mstd2;
    constexpr int KCEN = 3 * S + (S / 2 - 1);  // window centre: long offset 0, short offset 0
    orig = (KCEN & 1) ? xs[KCEN / 2].y : xs[KCEN / 2].x;
    // (x - mean) * inv_std as one FFMA2 per pair: x * inv_std - mean * inv_std.  The product term is rounded once
    // more than in the subtract-first form; its error (<= 2^-24 |x inv_std|) stays below the binary16 rounding of
    // the operand that follows even for the flattest admissible windows (variance >= 1.19e-7), where the output
    // contribution is scaled by the tiny stddev anyway.
    const float2 sc = make_float2(mstd2, mstd2), of = make_float2(-mstd0 * mstd2, -mstd0 * mstd2);
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      __half2 h[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 nv = __ffma2_rn(xs[kc * 4 + e], sc, of);
        h[e] = __floats2half2_rn(nv.x, nv.y);
      }
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h[0]);
      pk.y = *reinterpret_cast<uint32_t*>(&h[1]);
      pk.z = *reinterpret_cast<uint32_t*>(&h[2]);
      pk.w = *reinterpret_cast<uint32_t*>(&h[3]);
      *reinterpret_cast<uint4*>(abuf + kc * (128 * 16) + lt * 16) = pk;
    }
  };

  const long long tile_step = (long long)gridDim.x * kWG;

  struct TileCtx {
    float mstd0, mstd1, orig;
    int x0, y0, f;
  };
  // front half of a tile: staged window -> A operand (buffer `slot`) -> MMA of chunk 0 into TMEM slot `slot`
  // `ahead` always stands one tile beyond the tile whose front half runs (it is the tile being prefetched)
  TileWalk ahead((long long)blockIdx.x * kWG + wg, tile_step, A.tiles_x, A.tiles_y);
  // XT: staged window -> A operand in buffer `aslot`; no barrier, no MMA (the caller's next barrier publishes A)
  auto front_build = [&](int aslot, TileCtx& tc) {
    tc.f = ahead.f;
    tc.x0 = ahead.tix * kTileW;
    tc.y0 = ahead.tiy * kTileH;
    ahead.next();
    if constexpr (TMA) {
      mbar_wait(stage_bar, stage_phase);
      stage_phase ^= 1;
      const bool edge = tc.x0 - XO < 0 || tc.y0 - OY < 0 || tc.x0 - XO + SW > A.w || tc.y0 - OY + SH > A.h;
      if (edge) patch_clamp_to_edge(my_stage, SW, SW, SH, tc.x0 - XO, tc.y0 - OY, A.w, A.h, lt, 128, [&] { wg_barrier(wg); });
    } else {
      asm volatile("cp.async.wait_all;" ::: "memory");
      wg_barrier(wg);
    }
    build_a(my_a + aslot * kABytes, tc.mstd0, tc.mstd1, tc.orig);
    fence_proxy_async();
  };
  auto front = [&](long long tile, int slot, TileCtx& tc) {
    tc.f = ahead.f;
    tc.x0 = ahead.tix * kTileW;
    tc.y0 = ahead.tiy * kTileH;
    ahead.next();

    if constexpr (TMA) {
      mbar_wait(stage_bar, stage_phase);   // every thread waits: the box has landed and is visible
      stage_phase ^= 1;
      const bool edge = tc.x0 - XO < 0 || tc.y0 - OY < 0 || tc.x0 - XO + SW > A.w || tc.y0 - OY + SH > A.h;
      if (edge)   // warpgroup-uniform: replicate the border over the zero-filled texels
        patch_clamp_to_edge(my_stage, SW, SW, SH, tc.x0 - XO, tc.y0 - OY, A.w, A.h, lt, 128, [&] { wg_barrier(wg); });
    } else {
      asm volatile("cp.async.wait_all;" ::: "memory");
      wg_barrier(wg);  // the staged window of this tile is complete and visible
    }

    // ---- im2col + normalisation -> A operand ----------------------------------------------------
    build_a(my_a + slot * kABytes, tc.mstd0, tc.mstd1, tc.orig);
    fence_proxy_async();   // generic-proxy writes of A -> visible to the tensor core (async proxy)
    tc_fence_before();     // orders the previous tiles' tcgen05.ld before the new MMAs
    wg_barrier(wg);        // A complete; nobody reads the staging buffer any more
    if (lt == 0) {
      tc_fence_after();
      issue_chunk(0, slot);
    }
    prefetch_tile(tile + tile_step, ahead);
  };

  // back half: epilogue over the accumulator chunks of the tile (TMEM slot `slot`), then the store
  // XT: `more` = another tile follows (tile index nxt_tile); its A operand goes to buffer `nxt_aslot`, context to `nxt`
  auto back = [&](const TileCtx& tc, int slot, uint32_t parity, int aslot = 0, bool more = false, int nxt_aslot = 0,
                  TileCtx* nxt = nullptr, long long nxt_tile = 0) {
    const float mstd0 = tc.mstd0, mstd1 = tc.mstd1, orig = tc.orig;
    const int x0 = tc.x0, y0 = tc.y0, f = tc.f;
    const uint32_t d_lane_s = d_lane + (uint32_t)(slot * CN);
    // ---- epilogue -----------------------------------------------------------------------------
    float wsum = 0.f, vsum = 0.f;
    float2 wsum2 = make_float2(0.f, 0.f), vsum2 = make_float2(0.f, 0.f);
    if constexpr (PP) {
      // 16 neurons: registers 0..15 softmax logits, 16..31 their elliott inputs; one Newton reciprocal per 8 neurons
      auto math32 = [&](const uint32_t (&cur)[32]) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float2 s1[4], u[4], pk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i0 = 8 * g + 2 * k;
            s1[k].x = ex2_approx(__uint_as_float(cur[i0]));
            s1[k].y = ex2_approx(__uint_as_float(cur[i0 + 1]));
            const float2 t = make_float2(__uint_as_float(cur[16 + i0]), __uint_as_float(cur[16 + i0 + 1]));
            u[k].x = 1.0f + fabsf(t.x);
            u[k].y = 1.0f + fabsf(t.y);
            pk[k] = __fmul2_rn(s1[k], t);
          }
          const float2 n01 = __ffma2_rn(pk[1], u[0], __fmul2_rn(pk[0], u[1]));
          const float2 n23 = __ffma2_rn(pk[3], u[2], __fmul2_rn(pk[2], u[3]));
          const float2 d01 = __fmul2_rn(u[0], u[1]), d23 = __fmul2_rn(u[2], u[3]);
          const float2 num = __ffma2_rn(n23, d01, __fmul2_rn(n01, d23));
          const float2 den = __fmul2_rn(d01, d23);
          float2 rn;
          rn.x = __int_as_float((int)(0x7EF311C7u + 0x80000000u) - __float_as_int(den.x));
          rn.y = __int_as_float((int)(0x7EF311C7u + 0x80000000u) - __float_as_int(den.y));
          rn = __fmul2_rn(rn, __ffma2_rn(den, rn, make_float2(2.f, 2.f)));
          rn = __fmul2_rn(rn, __ffma2_rn(den, rn, make_float2(2.f, 2.f)));
          vsum2 = __ffma2_rn(num, rn, vsum2);   // accumulates -vsum
          wsum2 = __fadd2_rn(wsum2, __fadd2_rn(__fadd2_rn(s1[0], s1[1]), __fadd2_rn(s1[2], s1[3])));
        }
      };
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const uint32_t sl = (uint32_t)(c & 1);
        mbar_wait(my_mbar + 8u * sl, (full_par >> sl) & 1u);
        full_par ^= 1u << sl;
        tc_fence_after();
        uint32_t va[32], vb[32];
        [[maybe_unused]] uint32_t vc[32], vd[32];
        tmem_ld32_issue(d_lane + sl * (uint32_t)CN, va);
        tmem_ld32_issue(d_lane + sl * (uint32_t)CN + 32u, vb);
        if constexpr (CN == 128) {
          tmem_ld32_issue(d_lane + sl * (uint32_t)CN + 64u, vc);
          tmem_ld32_issue(d_lane + sl * (uint32_t)CN + 96u, vd);
        }
        tmem_ld_wait();
        // this warp's share of the slot is in registers
        tc_fence_before();
        if (c + 2 < NCH && (lt & 31) == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(free_bar + 8u * sl) : "memory");
        if (lt == 0 && c + 1 < NCH) {
          // MMA of the next chunk into the other slot: every warp loaded that slot's previous contents a chunk ago
          if (c >= 1) {
            mbar_wait(free_bar + 8u * (sl ^ 1u), (free_par >> (sl ^ 1u)) & 1u);
            free_par ^= 1u << (sl ^ 1u);
          }
          tc_fence_after();
          issue_chunk(c + 1, (int)(sl ^ 1u), aslot);
        }
        math32(va);
        math32(vb);
        if constexpr (CN == 128) {
          math32(vc);
          math32(vd);
        }
      }
    } else {
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      if constexpr (XT) {
        // the MMA of the last chunk is running: build the next tile's A operand under it
        if (c == NCH - 1 && more) front_build(nxt_aslot, *nxt);
      }
      if constexpr (DB) {
        mbar_wait(my_mbar + 8u * (uint32_t)slot, parity);
      } else {
        mbar_wait(my_mbar, phase);
        phase ^= 1;
      }
      tc_fence_after();
      uint32_t va[32], vb[32];
      tmem_ld32_issue(d_lane_s, va);
      // the accumulator is free once the LAST 32 columns of the chunk are in registers: then the MMA of the next chunk
      // (or, XT, of the next tile) is issued.  EARLY: that moment is moved from the top of the last block to the middle
      // of the second-to-last one (the last block's load is waited for after half of the previous block's arithmetic),
      // so that 1.5 blocks of arithmetic instead of 1 cover the MMA latency (the wait for the MMA at the top of a chunk
      // was 6.8 % of all warp samples in profiles/r02_nnedi3_*: the MMA was not done when the last block was).
#ifndef MPVP_X_NN_EARLY
#define MPVP_X_NN_EARLY 1
#endif
      constexpr int NB = CN / 32;
      constexpr bool EARLY = MPVP_X_NN_EARLY && EPI == 3 && NB >= 2;
#ifndef MPVP_X_NN_ARV
#define MPVP_X_NN_ARV 0   // chunk hand-over by mbarrier: the three non-issuing warps arrive and go on, only the issuing warp waits
#endif
      auto accumulator_free = [&]() {
        if (c + 1 < NCH) {
          tc_fence_before();
          if constexpr (MPVP_X_NN_ARV && !XT && !DB) {
            if ((lt & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(free_bar) : "memory");
            if (lt == 0) {
              mbar_wait(free_bar, free_par & 1u);
              tc_fence_after();
              issue_chunk(c + 1, 0, aslot);
            }
            free_par ^= 1u;
          } else {
          wg_barrier(wg);
          if (lt == 0) {
            tc_fence_after();
            issue_chunk(c + 1, 0, aslot);
          }
          }
        } else if (XT && more) {
          // last block of the tile: the accumulator is free and (same barrier) the next tile's A operand is complete
          tc_fence_before();
          wg_barrier(wg);
          if (lt == 0) {
            tc_fence_after();
            issue_chunk(0, 0, nxt_aslot);
          }
          prefetch_tile(nxt_tile + tile_step, ahead);   // the staging buffer was last read by front_build above
        }
      };
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        uint32_t (&cur)[32] = (i & 1) ? vb : va;
        uint32_t (&nxt)[32] = (i & 1) ? va : vb;
        // (tcgen05.wait::ld emits no instruction of its own: ptxas tracks the LDTM destination registers on the scoreboard,
        // so WHERE the wait stands does not change the SASS -- only where accumulator_free() stands does)
        if (!(EARLY && i == NB - 1)) tmem_ld_wait();   // EARLY: the last block was waited for inside the previous one
        if (i + 1 < NB) {
          tmem_ld32_issue(d_lane_s + (i + 1) * 32, nxt);
        } else if (!EARLY) {
          accumulator_free();
        }
        // registers 0..15: softmax logits of 16 neurons, 16..31: their elliott inputs
        if constexpr (EPI == 3) {
          // packed f32x2, ONE Newton reciprocal per 8 neurons: the four packed elliott terms p_k / u_k
          // (p = s1 * t, u = 1 + |t|) are put over the common denominator u0 u1 u2 u3
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float2 s1[4], u[4], pk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int i0 = 8 * g + 2 * k;
              s1[k].x = ex2_approx(__uint_as_float(cur[i0]));
              s1[k].y = ex2_approx(__uint_as_float(cur[i0 + 1]));
              const float2 t = make_float2(__uint_as_float(cur[16 + i0]), __uint_as_float(cur[16 + i0 + 1]));
              u[k].x = 1.0f + fabsf(t.x);   // scalar FADD with the |x| source modifier
              u[k].y = 1.0f + fabsf(t.y);
              pk[k] = __fmul2_rn(s1[k], t);
            }
            const float2 n01 = __ffma2_rn(pk[1], u[0], __fmul2_rn(pk[0], u[1]));
            const float2 n23 = __ffma2_rn(pk[3], u[2], __fmul2_rn(pk[2], u[3]));
            const float2 d01 = __fmul2_rn(u[0], u[1]), d23 = __fmul2_rn(u[2], u[3]);
            const float2 num = __ffma2_rn(n23, d01, __fmul2_rn(n01, d23));
            const float2 den = __fmul2_rn(d01, d23);
            // rn ~ -1/den: bit-trick seed with the sign bit set, two Newton steps rn <- rn * (2 + den * rn)
            float2 rn;
            rn.x = __int_as_float((int)(0x7EF311C7u + 0x80000000u) - __float_as_int(den.x));
            rn.y = __int_as_float((int)(0x7EF311C7u + 0x80000000u) - __float_as_int(den.y));
            rn = __fmul2_rn(rn, __ffma2_rn(den, rn, make_float2(2.f, 2.f)));
            rn = __fmul2_rn(rn, __ffma2_rn(den, rn, make_float2(2.f, 2.f)));
            vsum2 = __ffma2_rn(num, rn, vsum2);   // accumulates -vsum
            wsum2 = __fadd2_rn(wsum2, __fadd2_rn(__fadd2_rn(s1[0], s1[1]), __fadd2_rn(s1[2], s1[3])));
            if (EARLY && g == 0 && i == NB - 2) {
              tmem_ld_wait();        // the chunk's last 32 columns (issued at the top of this block) are in registers
              accumulator_free();
            }
          }
        } else if constexpr (EPI == 2) {
          // packed f32x2: two neurons per instruction on the FMA pipe
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float2 s1, t, a, r;
            s1.x = ex2_approx(__uint_as_float(cur[2 * e]));
            s1.y = ex2_approx(__uint_as_float(cur[2 * e + 1]));
            t.x = __uint_as_float(cur[16 + 2 * e]);
            t.y = __uint_as_float(cur[16 + 2 * e + 1]);
            a.x = fabsf(t.x);
            a.y = fabsf(t.y);
            const float2 nu = __ffma2_rn(a, make_float2(-1.f, -1.f), make_float2(-1.f, -1.f));  // -(1 + |t|)
            // seed for 1/u from the bits of -u: magic - (bits & 0x7fffffff) == (magic + 0x80000000) - bits
            r.x = __int_as_float((int)(0x7EF311C7u + 0x80000000u) - __float_as_int(nu.x));
            r.y = __int_as_float((int)(0x7EF311C7u + 0x80000000u) - __float_as_int(nu.y));
            r = __fmul2_rn(r, __ffma2_rn(nu, r, make_float2(2.f, 2.f)));
            r = __fmul2_rn(r, __ffma2_rn(nu, r, make_float2(2.f, 2.f)));
            wsum2 = __fadd2_rn(wsum2, s1);
            vsum2 = __ffma2_rn(__fmul2_rn(s1, t), r, vsum2);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float s1 = ex2_approx(__uint_as_float(cur[e]));
            const float t = __uint_as_float(cur[16 + e]);
            const float u = 1.0f + fabsf(t);
            const float r = (EPI == 0) ? rcp_approx(u) : rcp_newton(u);
            wsum += s1;
            vsum = fmaf(s1 * t, r, vsum);
          }
        }
      }
    }

    }   // !PP
    if constexpr (EPI == 2) {
      wsum = wsum2.x + wsum2.y;
      vsum = vsum2.x + vsum2.y;
    }
    if constexpr (EPI == 3) {
      wsum = wsum2.x + wsum2.y;
      vsum = -(vsum2.x + vsum2.y);
    }
    const int x = x0 + tx, y = y0 + ty;
    if (x < A.w && y < A.h) {
      const float pred = fminf(fmaxf(mstd0 + 5.0f * vsum / wsum * mstd1, 0.f), 1.f);
      const int64_t o = (int64_t)f * A.out_sn;
      if (DIR == 0) {
        store_px(A.out, o + (int64_t)(2 * y) * A.out_sy + x, orig, A.io.out_fmt, A.io.out_max);
        store_px(A.out, o + (int64_t)(2 * y + 1) * A.out_sy + x, pred, A.io.out_fmt, A.io.out_max);
      } else {
        store_px2(A.out, o + (int64_t)y * A.out_sy + 2 * x, orig, pred, A.io.out_fmt, A.io.out_max);
      }
    }
  };

  const long long first = (long long)blockIdx.x * kWG + wg;
  prefetch_tile(first, ahead);
  if constexpr (DB) {
    // two tiles in flight: front(t+1) is issued before back(t)
    if (first < A.total_tiles) {
      TileCtx cur, nxt;
      front(first, 0, cur);
      uint32_t it = 0;
      for (long long tile = first; tile < A.total_tiles; tile += tile_step, ++it) {
        const bool more = tile + tile_step < A.total_tiles;
        if (more) front(tile + tile_step, (int)((it + 1) & 1), nxt);
        back(cur, (int)(it & 1), (it >> 1) & 1);
        if (more) cur = nxt;
      }
    }
  } else if constexpr (XT) {
    if (first < A.total_tiles) {
      TileCtx cur, nxt;
      front(first, 0, cur);    // first tile: build, barrier, MMA of chunk 0, prefetch of the second tile's window
      uint32_t it = 0;
      for (long long tile = first; tile < A.total_tiles; tile += tile_step, ++it) {
        const bool more = tile + tile_step < A.total_tiles;
        back(cur, 0, 0, (int)(it & 1), more, (int)((it + 1) & 1), &nxt, tile + tile_step);
        if (more) cur = nxt;
      }
    }
  } else {
    for (long long tile = first; tile < A.total_tiles; tile += tile_step) {
      TileCtx tc;
      front(tile, 0, tc);
      back(tc, 0, 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Epilogue variants (A/B switch MPVP_NNEDI3_EPI): 0 = reciprocal on the MUFU pipe, 1 = scalar Newton
// reciprocal on the FMA pipe, 2 = Newton in packed f32x2 arithmetic (two neurons per instruction), 3 (default) =
// packed, with one Newton reciprocal shared by 8 neurons (common denominator).
static int epi_mode() {
  static const int v = [] {
    const char* e = getenv("MPVP_NNEDI3_EPI");
    return (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 3;
  }();
  return v;
}

template <int S, int DIR, int NNS>
int launch_tc(const NnTcArgs& a0, int device, cudaStream_t stream) {
  MPVP_REQUIRE(a0.group == 16, "NNEDI3 weights were packed with %d neurons per accumulator block, the kernel expects 16", a0.group);
  constexpr int K = 8 * S, KX = K + 16, N = 2 * NNS;
  constexpr int HX = DIR == 0 ? 8 : S, HY = DIR == 0 ? S : 8;
  constexpr int SH = kTileH + HY - 1;
  constexpr int NA = ((N <= 64) || MPVP_X_NN_XT) ? 2 : 1;  // A buffers per warpgroup (two tiles in flight / cross-tile pipelining)
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  const bool use_tma = a0.io.in_fmt == MPVP_FMT_F32 && make_plane_tmap(&tmap, a0.in, 4, a0.w, a0.h, a0.n, a0.in_sy, a0.in_sn, 40, SH);
  const int SW = use_tma ? 40 : (kTileW + HX - 1);
  const int STG = (SW * SH + 31) & ~31;
  size_t smem = (size_t)KX * N * 2 + (size_t)kWG * NA * KX * 128 * 2 + sizeof(float) * kWG * STG + 8 * (5 * kWG + 1) + 16;
  // one CTA per SM: the kernel allocates all 512 TMEM columns
  if (smem < 120 * 1024) smem = 120 * 1024;
  NnTcArgs a = a0;
  a.tiles_x = (a.w + kTileW - 1) / kTileW;
  a.tiles_y = (a.h + kTileH - 1) / kTileH;
  a.total_tiles = (long long)a.tiles_x * a.tiles_y * a.n;
  MPVP_REQUIRE(a.total_tiles < (1LL << 31), "batch too large: %lld tiles (limit 2^31)", a.total_tiles);
  const int mode = epi_mode();
  auto kern = use_tma ? nnedi3_tc_kernel<S, DIR, NNS, 3, true> : nnedi3_tc_kernel<S, DIR, NNS, 3, false>;
  if (mode != 3) {   // A/B epilogue variants run on the plain-load staging path
    kern = mode == 0 ? nnedi3_tc_kernel<S, DIR, NNS, 0, false>
           : mode == 1 ? nnedi3_tc_kernel<S, DIR, NNS, 1, false>
                       : nnedi3_tc_kernel<S, DIR, NNS, 2, false>;
  }
  MPVP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = sm_count(device);
  const long long need = (a.total_tiles + kWG - 1) / kWG;
  if (grid > need) grid = need;
  grid = cap_grid(grid);
  if (grid < 1) return MPVP_OK;
  kern<<<(unsigned)grid, kThreads, smem, stream>>>(a, tmap);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MPVP_CUDA_OK(cudaGetLastError());
  return MPVP_OK;
}

template <int S, int DIR>
int dispatch_nns(const NnTcArgs& a, int nns, int device, cudaStream_t st) {
  switch (nns) {
    case 16: return launch_tc<S, DIR, 16>(a, device, st);
    case 32: return launch_tc<S, DIR, 32>(a, device, st);
    case 64: return launch_tc<S, DIR, 64>(a, device, st);
    case 128: return launch_tc<S, DIR, 128>(a, device, st);
    case 256: return launch_tc<S, DIR, 256>(a, device, st);
  }
  set_error("nns %d unsupported", nns);
  return MPVP_E_UNSUPPORTED;
}

}  // namespace

// B row order the kernel expects: blocks of 32 accumulator columns = 16 neurons (16 logits, then their 16 elliott inputs)
int nnedi3_group_size(int) { return 16; }

int nnedi3_tc(const mpvp_weights* nn, int direction, const void* in, void* out, int n, int h, int w,
              int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n, int64_t out_stride_y, const IoFmt& io,
              cudaStream_t st) {
  NnTcArgs a{};
  a.io = io;
  a.in = in; a.out = out; a.b_packed = nn->nn_b; a.bias = nn->nn_bias; a.group = nn->nn_group;
  a.n = n; a.h = h; a.w = w;
  a.in_sn = in_stride_n; a.in_sy = in_stride_y; a.out_sn = out_stride_n; a.out_sy = out_stride_y;
  if (nn->win_short == 4)
    return direction == 0 ? dispatch_nns<4, 0>(a, nn->nns, nn->device, st) : dispatch_nns<4, 1>(a, nn->nns, nn->device, st);
  return direction == 0 ? dispatch_nns<6, 0>(a, nn->nns, nn->device, st) : dispatch_nns<6, 1>(a, nn->nns, nn->device, st);
}

}  // namespace mpvp
