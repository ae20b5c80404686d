// NADE / MultiNADE teacher-forced log-likelihood on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), segment-row form.
// Same arithmetic as nade.cu::nade_fwd_kernel (reference multinn/models/common/nade.py:155-229, :310-329,
// utils/auxiliary.py:9-11; per-track loop of generators/rnn_multinade.py:258-290), other schedule:
//
//   a_{i+1} = a_i + v_i w_enc[i] only changes at set target bits, so a (row, track) pair with K set bits among dims
//   0..D-2 has K+1 distinct hidden vectors h_s = sigmoid(a_s); stack them as SEGMENT ROWS. A tile = 128 segment rows
//   (whole source rows, packed greedily) = the 128 TMEM lanes of one accumulator:
//       Lt[128 x 96] = H[128 x 256] . W_dec^T[256 x 96]          (tcgen05.mma kind::f16, bf16 operands, fp32 in TMEM)
//   entry (r, i) is the logit of dim i iff segment row r owns dim i. fp32 accuracy by operand splitting:
//   h = h1 + h2, w = w1 + w2 (bf16 each, 16 mantissa bits together); h1.w1 goes to the main accumulator, h1.w2 + h2.w1
//   to a second one (dropped terms ~2^-17 |h w|; logits agree with the fp32 SIMT kernel to ~1e-5 absolute).
//
// One CTA = one track (its split W_dec is resident in shared memory in the UMMA canonical K-major SWIZZLE_64B layout) and a
// contiguous range of source rows; 512 threads:
//   warp 0      TMA: the k-slice W_enc[m][:, 32 kb .. 32 kb + 31] of each k-block -> 3-slot ring (mbarrier complete_tx)
//   warp 1      MMA issuer (one thread): 6 MMAs per k-block (2 K=16 steps x 3 products), commits free the A stage /
//               publish the accumulator (double buffered: 2 x (main + aux) = 512 TMEM columns)
//   warps 4-7   epilogue: tcgen05.ld (lane = segment row) -> Lt staging tile in shared memory -> balanced pass, one warp
//               per SOURCE row, lanes over dims: + b_dec, sigmoid, BCE with the 1e-6 eps, d b_dec, cond_p, NLL warp-sum
//   warps 8-15  producers, two phases per k-block: (1) task = (source row, 4 hidden units): the prefixes a_r = b_enc +
//               sum of w_enc[bit] over the bits that opened the row's segments, staged as plain fp32 in the A stage's own
//               memory; (2) item = (segment row, 4 units), 4 per thread: sigmoid, split into bf16 h1 / h2, swizzled 8-byte
//               stores of both A tiles; warp 8 also packs the next tile (warp scan of 1 + popcount over up to 128 rows)
// Tiles never straddle source rows, so the NLL reduction stays inside a warp and nothing is atomically accumulated.
#include <cuda.h>

#include <cstdlib>

#include "multinn_b200.h"
#include "tc_common.cuh"

namespace mnn {
namespace ntc {

using namespace mnn::tc;

constexpr int kThreads = 512;
constexpr int kProdWarps = 8, kProdThreads = kProdWarps * 32;
constexpr int kEpiWarps = 4, kEpiThreads = kEpiWarps * 32;
constexpr int DP = 96;                 // MMA N: dims padded to a multiple of 16
constexpr int LTLD = 97;               // row stride of the Lt staging tile (odd: conflict-free column walks)
constexpr int kWencSlots = 3, kASlots = 2, kInfoSlots = 4;
constexpr int B_TILE = DP * 64;        // one [96 x 32] bf16 K-major SWIZZLE_64B tile
constexpr int A_TILE = 128 * 64;       // one [128 x 32] bf16 tile
constexpr int WENC_SLOT = DP * 128;    // [<=96 rows][32] fp32
constexpr int kMaxIter = 4;            // task iterations per producer thread: 128 rows x 8 quads / 256 threads

struct TileInfo {
  uint32_t bits[128][4];   // target masks of the tile's source rows (bit i = v_i)
  uint8_t base[128];       // first segment row of source row j
  uint8_t segsrc[128];     // source row j of segment row r
  uint8_t segpos[128];     // the set bit (dim) that opened segment row r (unused for a row's first segment)
  int row0, cnt, nseg, pad;
};

struct Args {
  const uint32_t* bits; const float* fc; long long ld; int enc_col0, dec_col0;
  const float* w_dec;
  float* nll; float* cond_p; float* dfc; float gscale;
  int N, M, D, H;
  long long tstride;
};

template <int KB>
struct Smem {
  static constexpr int OFF_B1 = 0;
  static constexpr int OFF_B2 = KB * B_TILE;
  static constexpr int OFF_WENC = 2 * KB * B_TILE;
  static constexpr int OFF_A = OFF_WENC + kWencSlots * WENC_SLOT;
  static constexpr int OFF_LT = OFF_A + kASlots * 2 * A_TILE;
  static constexpr int OFF_INFO = OFF_LT + 128 * LTLD * 4;
  static constexpr int TOTAL = OFF_INFO + kInfoSlots * (int)sizeof(TileInfo) + 1024;
};

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

// byte offset of elements (row, k0 .. k0+3), k0 % 4 == 0, inside a [rows x 32] bf16 K-major SWIZZLE_64B tile
__device__ __forceinline__ uint32_t sw64_off(int row, int k0) {
  const int r8 = row & 7;
  return (uint32_t)((row >> 3) * 512 + r8 * 64 + (((k0 >> 3) ^ (r8 >> 1)) << 4) + ((k0 & 7) << 1));
}

// x = hi + lo with hi = bf16(x) (round to nearest) and lo = bf16(x - hi): 8 bytes of each for 4 consecutive elements
__device__ __forceinline__ void split_bf16x4(float4 x, uint2& hi, uint2& lo) {
  hi = pack_bf16x4(x.x, x.y, x.z, x.w);
  const float rx = x.x - __uint_as_float(hi.x << 16), ry = x.y - __uint_as_float(hi.x & 0xffff0000u);
  const float rz = x.z - __uint_as_float(hi.y << 16), rw = x.w - __uint_as_float(hi.y & 0xffff0000u);
  lo = pack_bf16x4(rx, ry, rz, rw);
}

__device__ __forceinline__ float4 sig4(float4 a) {
  return make_float4(sigmoid_mufu(a.x), sigmoid_mufu(a.y), sigmoid_mufu(a.z), sigmoid_mufu(a.w));
}

template <int KB>
__global__ void __launch_bounds__(kThreads, 1)
nade_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_wenc, const Args p) {
  using S = Smem<KB>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kWencSlots + 2 * kASlots + 4 + 2 * kInfoSlots];
  __shared__ uint32_t tmem_base_s;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem0 - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = p.D, H = KB * 32;

  const uint32_t bar_wfull = smem_u32(&bars[0]);                        // [3] W_enc slice landed
  const uint32_t bar_wempty = bar_wfull + 8 * kWencSlots;               // [3] producers done with the slice
  const uint32_t bar_afull = bar_wempty + 8 * kWencSlots;               // [2] A tiles written
  const uint32_t bar_aempty = bar_afull + 8 * kASlots;                  // [2] MMAs have read the A tiles
  const uint32_t bar_tfull = bar_aempty + 8 * kASlots;                  // [2] accumulator complete
  const uint32_t bar_tempty = bar_tfull + 16;                           // [2] accumulator drained
  const uint32_t bar_ifull = bar_tempty + 16;                           // [4] tile descriptor written
  const uint32_t bar_iempty = bar_ifull + 8 * kInfoSlots;               // [4] epilogue done with the descriptor

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWencSlots; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, kProdWarps); }
    for (int s = 0; s < kASlots; ++s) { mbar_init(bar_afull + 8 * s, kProdWarps); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, kEpiWarps); }
    for (int s = 0; s < kInfoSlots; ++s) { mbar_init(bar_ifull + 8 * s, 1); mbar_init(bar_iempty + 8 * s, kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // this CTA's track and source-row range
  const int m = blockIdx.x % p.M, ci = blockIdx.x / p.M;
  const int nc = (gridDim.x - m + p.M - 1) / p.M;
  const int row_begin = (int)((long long)p.N * ci / nc), row_end = (int)((long long)p.N * (ci + 1) / nc);

  // resident B operand: W_dec[m] split into bf16 w1 + w2, [KB] tiles of [96 dims x 32 hidden] each (rows >= D zero)
  {
    const float* wd = p.w_dec + (size_t)m * D * H;
    for (int idx = threadIdx.x; idx < DP * (H / 4); idx += kThreads) {
      const int i = idx / (H / 4), k = (idx - i * (H / 4)) * 4;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < D) w = __ldg(reinterpret_cast<const float4*>(wd + (size_t)i * H + k));
      uint2 hi, lo;
      split_bf16x4(w, hi, lo);
      const uint32_t off = (uint32_t)((k >> 5) * B_TILE) + sw64_off(i, k & 31);
      *reinterpret_cast<uint2*>(smem + S::OFF_B1 + off) = hi;
      *reinterpret_cast<uint2*>(smem + S::OFF_B2 + off) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  TileInfo* infos = reinterpret_cast<TileInfo*>(smem + S::OFF_INFO);

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA: W_enc k-slices
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int t = 0;; ++t) {
        mbar_wait(bar_ifull + 8 * (t & 3), (t >> 2) & 1);
        if (infos[t & 3].cnt == 0) break;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_wempty + 8 * ws, wphase ^ 1);
          mbar_expect_tx(bar_wfull + 8 * ws, (uint32_t)D * 128u);
          tma_load_2d(smem0 + S::OFF_WENC + ws * WENC_SLOT, &map_wenc, bar_wfull + 8 * ws, kb * 32, m * D);
          if (++ws == kWencSlots) { ws = 0; wphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(DP, false, false, 128);
      int as = 0, acc = 0;
      uint32_t aphase = 0, acc_phase = 0;
      for (int t = 0;; ++t) {
        mbar_wait(bar_ifull + 8 * (t & 3), (t >> 2) & 1);
        if (infos[t & 3].cnt == 0) break;
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(acc * 256), d_aux = d_main + 128;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_afull + 8 * as, aphase);
          tc_fence_after();
          const uint32_t a1 = smem0 + S::OFF_A + as * 2 * A_TILE, a2 = a1 + A_TILE;
          const uint32_t b1 = smem0 + S::OFF_B1 + kb * B_TILE, b2 = smem0 + S::OFF_B2 + kb * B_TILE;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t da1 = smem_desc(a1 + j * 32, 16, 512, 4), da2 = smem_desc(a2 + j * 32, 16, 512, 4);
            const uint64_t db1 = smem_desc(b1 + j * 32, 16, 512, 4), db2 = smem_desc(b2 + j * 32, 16, 512, 4);
            const uint32_t first = (kb > 0 || j > 0) ? 1u : 0u;
            umma_bf16(d_main, da1, db1, idesc, first);
            umma_bf16(d_aux, da1, db2, idesc, first);
            umma_bf16(d_aux, da2, db1, idesc, 1u);
          }
          umma_commit(bar_aempty + 8 * as);
          if (++as == kASlots) { as = 0; aphase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ---------------------------------------------------------------- producers
    const int tid = threadIdx.x - 8 * 32;
    const uint32_t last_word_mask = ((D - 1) & 31) ? ((1u << ((D - 1) & 31)) - 1u) : 0u;   // dims 0..D-2 open segments
    const int last_word = (D - 1) >> 5;
    int next_row = row_begin;

    // pack the next tile: whole source rows, at most 128 segment rows (1 + popcount of the bits that open a segment)
    auto build_info = [&](int t) {
      TileInfo& ti = infos[t & 3];
      if (t >= kInfoSlots) mbar_wait(bar_iempty + 8 * (t & 3), ((t >> 2) & 1) ^ 1);
      int carry = 0, cnt = 0;
      const uint32_t* bits = p.bits + (size_t)m * p.tstride * 4;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        const int row = next_row + g * 32 + lane;
        uint4 b = make_uint4(0u, 0u, 0u, 0u);
        int ns = 1000;                                  // rows past the range never fit
        if (row < row_end) {
          b = __ldg(reinterpret_cast<const uint4*>(bits + (size_t)row * 4));
          uint32_t w[4] = {b.x, b.y, b.z, b.w};
          ns = 1;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t mk = q < last_word ? 0xffffffffu : (q == last_word ? last_word_mask : 0u);
            ns += __popc(w[q] & mk);
          }
        }
        int incl = ns;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        incl += carry;
        const bool fits = incl <= 128;
        const unsigned fm = __ballot_sync(0xffffffffu, fits);
        const int nfit = __popc(fm);                    // monotone: the fitting rows are lanes 0..nfit-1
        if (fits) {
          const int j = g * 32 + lane;
          ti.bits[j][0] = b.x; ti.bits[j][1] = b.y; ti.bits[j][2] = b.z; ti.bits[j][3] = b.w;
          int r = incl - ns;
          ti.base[j] = (uint8_t)r;
          ti.segsrc[r] = (uint8_t)j;
          uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t x = w[q] & (q < last_word ? 0xffffffffu : (q == last_word ? last_word_mask : 0u));
            while (x) {
              const int i = q * 32 + __ffs(x) - 1;
              x &= x - 1;
              ++r;
              ti.segsrc[r] = (uint8_t)j;
              ti.segpos[r] = (uint8_t)i;
            }
          }
        }
        cnt += nfit;
        if (nfit > 0) carry = __shfl_sync(0xffffffffu, incl, nfit - 1);
        if (nfit < 32) break;
      }
      if (lane == 0) { ti.row0 = next_row; ti.cnt = cnt; ti.nseg = carry; }
      next_row += cnt;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ifull + 8 * (t & 3));
    };

    // task = (source row j, unit quad q) of the k-block: its b_enc quad, prefetched one k-block ahead
    auto load_be = [&](const TileInfo& ti, int kb, float4 (&dst)[kMaxIter]) {
      const int ntask = ti.cnt * 8;
#pragma unroll
      for (int it = 0; it < kMaxIter; ++it) {
        const int task = tid + it * kProdThreads;
        if (task < ntask) {
          const int j = task >> 3, q = task & 7;
          dst[it] = __ldg(reinterpret_cast<const float4*>(p.fc + (size_t)(ti.row0 + j) * p.ld + p.enc_col0 + m * H +
                                                          kb * 32 + q * 4));
        }
      }
    };

    if (warp == 8) build_info(0);
    float4 cur[kMaxIter], nxt[kMaxIter];
    mbar_wait(bar_ifull, 0);
    if (infos[0].cnt > 0) load_be(infos[0], 0, cur);
    int ws = 0, as = 0;
    uint32_t wphase = 0, aphase = 0;
    for (int t = 0;; ++t) {
      const TileInfo& ti = infos[t & 3];
      if (t > 0) mbar_wait(bar_ifull + 8 * (t & 3), (t >> 2) & 1);
      if (ti.cnt == 0) break;
      if (warp == 8) build_info(t + 1);
      for (int kb = 0; kb < KB; ++kb) {
        // prefetch the b_enc slices of the next k-block (or of the next tile's first one)
        if (kb + 1 < KB) {
          load_be(ti, kb + 1, nxt);
        } else {
          mbar_wait(bar_ifull + 8 * ((t + 1) & 3), ((t + 1) >> 2) & 1);
          if (infos[(t + 1) & 3].cnt > 0) load_be(infos[(t + 1) & 3], 0, nxt);
        }
        mbar_wait(bar_wfull + 8 * ws, wphase);
        mbar_wait(bar_aempty + 8 * as, aphase ^ 1);
        const float* wenc = reinterpret_cast<const float*>(smem + S::OFF_WENC + ws * WENC_SLOT);
        uint8_t* a1 = smem + S::OFF_A + as * 2 * A_TILE;
        uint8_t* a2 = a1 + A_TILE;
        // phase 1 -- prefixes: task = (source row, unit quad) walks the bits that opened the row's segments with adds only
        // and leaves a_r of every segment row in the stage's own memory as a plain [128][32] fp32 tile (the serial part:
        // ~5 dependent adds at 5 % density). Measured alternatives: sigmoid + split inside this chain
        // 12.4 ms, independent items that redo the prefix per segment row 13.4 ms, this two-phase form 10.9 ms at C5.
        float4* stage = reinterpret_cast<float4*>(a1);
        const int ntask = ti.cnt * 8;
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
          const int task = tid + it * kProdThreads;
          if (task < ntask) {
            const int j = task >> 3, q = task & 7;
            const int r0 = ti.base[j], r1 = (j + 1 < ti.cnt) ? ti.base[j + 1] : ti.nseg;
            float4 a = cur[it];
            stage[r0 * 8 + q] = a;
            for (int r = r0 + 1; r < r1; ++r) {
              const float4 we = *reinterpret_cast<const float4*>(wenc + ti.segpos[r] * 32 + q * 4);
              a.x += we.x; a.y += we.y; a.z += we.z; a.w += we.w;
              stage[r * 8 + q] = a;
            }
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        // phase 2 -- every (segment row, quad) item is independent: 4 per thread, balanced whatever the rows look like
        const int nitem = ti.nseg * 8;
        float4 av[kMaxIter];
#pragma unroll
        for (int e = 0; e < kMaxIter; ++e) {
          const int idx = tid + e * kProdThreads;
          if (idx < nitem) av[e] = stage[idx];
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");     // all prefixes read before the tiles overwrite them
#pragma unroll
        for (int e = 0; e < kMaxIter; ++e) {
          const int idx = tid + e * kProdThreads;
          if (idx < nitem) {
            uint2 hi, lo;
            split_bf16x4(sig4(av[e]), hi, lo);
            const uint32_t off = sw64_off(idx >> 3, (idx & 7) * 4);
            *reinterpret_cast<uint2*>(a1 + off) = hi;
            *reinterpret_cast<uint2*>(a2 + off) = lo;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_afull + 8 * as);
          mbar_arrive(bar_wempty + 8 * ws);
        }
        if (++ws == kWencSlots) { ws = 0; wphase ^= 1; }
        if (++as == kASlots) { as = 0; aphase ^= 1; }
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) cur[it] = nxt[it];
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int q = warp - 4;
    float* lt = reinterpret_cast<float*>(smem + S::OFF_LT);
    int acc = 0;
    uint32_t acc_phase = 0;
    const int dec_col = p.dec_col0 + m * D;
    for (int t = 0;; ++t) {
      const TileInfo& ti = infos[t & 3];
      mbar_wait(bar_ifull + 8 * (t & 3), (t >> 2) & 1);
      if (ti.cnt == 0) break;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      asm volatile("bar.sync 1, 128;" ::: "memory");    // the previous tile's balanced pass has finished reading lt
      {
        const int r = q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
#pragma unroll 1
        for (int c = 0; c < DP / 32; ++c) {
          float v[32], vx[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld32(taddr + 128 + c * 32, vx);
#pragma unroll
          for (int i = 0; i < 32; ++i) lt[r * LTLD + c * 32 + i] = v[i] + vx[i];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      asm volatile("bar.sync 1, 128;" ::: "memory");    // lt complete
      // balanced pass: one warp per source row, lanes over dims i = lane + 32 it
      float bd_next[3] = {0.f, 0.f, 0.f};
      if (q < ti.cnt) {
        const float* fr = p.fc + (size_t)(ti.row0 + q) * p.ld + dec_col;
#pragma unroll
        for (int it = 0; it < 3; ++it)
          if (lane + 32 * it < D) bd_next[it] = __ldg(fr + lane + 32 * it);
      }
      for (int j = q; j < ti.cnt; j += kEpiWarps) {
        const int n = ti.row0 + j;
        const uint32_t w0 = ti.bits[j][0], w1 = ti.bits[j][1], w2 = ti.bits[j][2];
        const int base = ti.base[j];
        const uint32_t below = (1u << lane) - 1u;
        float bd[3] = {bd_next[0], bd_next[1], bd_next[2]};
        if (j + kEpiWarps < ti.cnt) {                       // next row's decoder biases: in flight under this row's math
          const float* fr = p.fc + (size_t)(n + kEpiWarps) * p.ld + dec_col;
#pragma unroll
          for (int it = 0; it < 3; ++it)
            if (lane + 32 * it < D) bd_next[it] = __ldg(fr + lane + 32 * it);
        }
        float nll_acc = 0.f;
#pragma unroll
        for (int it = 0; it < 3; ++it) {
          const int i = lane + 32 * it;
          if (i < D) {
            const uint32_t w = it == 0 ? w0 : (it == 1 ? w1 : w2);
            const int pre = it == 0 ? 0 : (it == 1 ? __popc(w0) : __popc(w0) + __popc(w1));
            const int s = pre + __popc(w & below);
            const bool v = (w >> lane) & 1u;
            const float l = lt[(base + s) * LTLD + i] + bd[it];
            const float pr = sigmoid_acc(l);
            const float qq = 1.0f - pr;
            nll_acc -= v ? logf(kSafeLogEps + pr) : logf(kSafeLogEps + qq);
            if (p.cond_p) p.cond_p[((size_t)m * p.tstride + n) * D + i] = pr;
            if (p.dfc) {
              const float pq = pr * qq;
              const float dl = v ? -pq / (kSafeLogEps + pr) : pq / (kSafeLogEps + qq);
              p.dfc[(size_t)n * p.ld + dec_col + i] = p.gscale * dl;
            }
          }
        }
        nll_acc = warp_sum(nll_acc);
        if (lane == 0) p.nll[(size_t)m * p.tstride + n] = nll_acc;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_iempty + 8 * (t & 3));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace ntc
}  // namespace mnn

using namespace mnn;

// 0 = tensor-core kernel where the shape fits, 1 = SIMT kernel (nade.cu) always. Default: 1 -- measured on a B200 at C5
// (profiles/r2_nade_tc_fwd.md) the SIMT segment kernel is still the faster forward (7.4 ms vs 11-13 ms): MNN_NADE_MODE=0
// or mnn_set_nade_mode(0) selects the tensor-core kernel.
static int g_nade_mode = -1;
extern "C" int mnn_set_nade_mode(int mode) {
  MNN_REQUIRE(mode == 0 || mode == 1, MNN_ERR_ARG, "set_nade_mode: 0 (tensor cores) or 1 (SIMT)");
  g_nade_mode = mode;
  return MNN_OK;
}

int mnn_nade_tc_wanted(int D, int H, long long ld, const float* w_enc) {
  if (g_nade_mode < 0) {
    const char* e = getenv("MNN_NADE_MODE");
    g_nade_mode = (e && e[0] == '0') ? 0 : 1;
  }
  if (g_nade_mode == 1) return 0;
  return (H == 128 || H == 256) && D >= 2 && D <= 96 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(w_enc) & 15) == 0;
}

int mnn_nade_tc_fwd(const uint32_t* bits, const float* fc, long long ld, int enc_col0, int dec_col0, const float* w_enc,
                    const float* w_dec, float* nll, float* cond_p, float* dfc, float gscale, int N, int M, int D, int H,
                    long long tstride, int sms, cudaStream_t stream) {
  CUtensorMap map;
  int rc = mnn_tc_make_map_plain(w_enc, H, H, (long long)M * D, 32, D, &map);
  if (rc) return rc;
  ntc::Args a{bits, fc, ld, enc_col0, dec_col0, w_dec, nll, cond_p, dfc, gscale, N, M, D, H, tstride};
  // a CTA should see a few tiles at least: ~25 source rows per tile at 5 % density
  int grid = sms;
  const int need = M * ((N + 63) / 64);
  if (grid > need) grid = need;
  if (grid < M) grid = M;
  if (H == 256) {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(ntc::nade_tc_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ntc::Smem<8>::TOTAL); set = true; }
    ntc::nade_tc_fwd_kernel<8><<<grid, ntc::kThreads, ntc::Smem<8>::TOTAL, stream>>>(map, a);
  } else {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(ntc::nade_tc_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ntc::Smem<4>::TOTAL); set = true; }
    ntc::nade_tc_fwd_kernel<4><<<grid, ntc::kThreads, ntc::Smem<4>::TOTAL, stream>>>(map, a);
  }
  return mnn_check_launch("nade_logprob_fwd(tc)");
}
