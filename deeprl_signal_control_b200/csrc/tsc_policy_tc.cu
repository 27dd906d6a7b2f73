// tsc_policy_tc.cu — fused per-control-step policy forward on the 5th-gen tensor cores (sm_100a).
//
// One persistent CTA (256 threads, 1 per SM, ~223 KB smem) walks (unit, 128-replica tile) work items:
//   1. SIMT fc front end (agents/policies.py:191-201): relu(fc) of the observation slice, written as
//      bf16 straight into the A-operand tile in shared memory (UMMA K-major, no swizzle), followed by
//      the previous hidden state h (masked by the pre-decision done flag, agents/utils.py:104-105);
//   2. one elected thread issues K/16 `tcgen05.mma.cta_group::1.kind::f16` (M=128, N=256, bf16 x bf16
//      -> fp32) against the packed [Wx;Wh] operand that stays resident in shared memory; the
//      accumulator lives in TMEM (256 columns); completion via tcgen05.commit -> mbarrier;
//   3. epilogue: every thread owns one replica row (TMEM lane) and 32 hidden units: tcgen05.ld of the
//      four gate pre-activations, bias, LSTM cell (agents/utils.py:106-113), state write-back, head
//      dot products (agents/policies.py:18-26), then softmax / value / categorical sample
//      (utils.py:155-157) for the row.
// Replaces per step: fc_embed_kernel + library GEMM + lstm_seq_fwd_kernel(T=1) + heads_kernel.
//
// Operand layouts (bytes, 16-byte "core rows", no swizzle; cute canonical ((8,n),2):((1,SBO),LBO)):
//   A tile  [KC][128 rows][16 B]   : LBO = 2048 (next K chunk of 8 bf16), SBO = 128 (next 8 rows)
//   B tile  [KC][256 cols][16 B]   : LBO = 4096,                          SBO = 128
#include <cuda.h>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/tsc_learn.h"

int tsc_set_error(const std::string& m);
#define PCK(call)                                                                  \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) return tsc_set_error(std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

struct DDimsTC {   // mirror of DDims in tsc_learn.cu (kept in sync by tscl_handle)
  int A, n_obs, max_na, fw, ff, ft, h, dx;
  const int32_t *obs_off, *n_wave, *n_wait, *n_fp, *n_a;
  const int64_t *off_fcw_w, *off_fcw_b, *off_fcf_w, *off_fcf_b, *off_fct_w, *off_fct_b;
  int64_t off_wx, off_wh, off_bl, off_wo, off_bo, n_params;
  int kw, ones_slot;
};
const DDimsTC* tscl_dims_of(tscl_handle* h);   // defined in tsc_learn.cu
int tscl_device_of(tscl_handle* h);

#define TC_M 128
#define TC_N 256
#define TC_H 64
#define TC_THREADS 256
#define TC_STAGE_ROWS 16
#define TC_KW 32
#define TC_KF 16
#define TC_KT 16
#define TC_KTOT (TC_KW + TC_KF + TC_KT)

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);   // version = 1 (Blackwell), no swizzle
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// MUFU tanh (tanh.approx.f32, max rel. error ~2^-11: below the bf16 operand rounding of this path)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ uint32_t pmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
  return h;
}

// ---------------------------------------------------------------------------------------------------
// per-unit stride of the packed image: main [KC][256][8] followed by the fc image [8][dx][8]
__host__ __device__ inline int64_t wp_stride(int dx) { return (int64_t)((dx + TC_H) / 8) * TC_N * 8 + (int64_t)8 * dx * 8; }

// pack [Wx;Wh] of every unit into the bf16 UMMA B-operand image  Wp[u][kc][n][8]
__global__ void pack_wxh_kernel(const DDimsTC d, const float* __restrict__ P, __nv_bfloat16* __restrict__ Wp) {
  const int u = blockIdx.y;
  const int K = d.dx + TC_H, KC = K / 8;
  const float* Wx = P + d.off_wx + (int64_t)u * d.dx * TC_N;
  const float* Wh = P + d.off_wh + (int64_t)u * TC_H * TC_N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < KC * TC_N; i += gridDim.x * blockDim.x) {
    const int kc = i / TC_N, n = i - kc * TC_N;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      const float w = k < d.dx ? Wx[(int64_t)k * TC_N + n] : Wh[(int64_t)(k - d.dx) * TC_N + n];
      v[e] = __float2bfloat16_rn(w);
    }
    *reinterpret_cast<uint4*>(Wp + (int64_t)u * wp_stride(d.dx) + ((int64_t)kc * TC_N + n) * 8) = *reinterpret_cast<const uint4*>(v);
  }
}


// block-diagonal fc weights as UMMA B operand: rows k = {wave 0..31 | fp 32..47 | wait 48..63}, cols = X columns
__global__ void pack_fc_kernel(const DDimsTC d, const float* __restrict__ P, __nv_bfloat16* __restrict__ Wp) {
  const int u = blockIdx.y, ag = u >> 1;
  const int nw = d.n_wave[ag], nt = d.n_wait[ag], nf = d.ff > 0 ? d.n_fp[ag] : 0;
  __nv_bfloat16* W0 = Wp + (int64_t)u * wp_stride(d.dx) + (int64_t)((d.dx + TC_H) / 8) * TC_N * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 8 * d.dx; i += gridDim.x * blockDim.x) {
    const int kc = i / d.dx, n = i - kc * d.dx;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      float w = 0.f;
      if (n < d.fw) { if (k < d.kw && k < nw) w = P[d.off_fcw_w[u] + (int64_t)k * d.fw + n]; }
      else if (n < d.fw + d.ff) { const int kk = k - d.kw; if (kk >= 0 && kk < TC_KF && kk < nf) w = P[d.off_fcf_w[u] + (int64_t)kk * d.ff + (n - d.fw)]; }
      else { const int kk = k - d.kw - TC_KF; if (kk >= 0 && kk < nt) w = P[d.off_fct_w[u] + (int64_t)kk * d.ft + (n - d.fw - d.ff)]; }
      v[e] = __float2bfloat16_rn(w);
    }
    *reinterpret_cast<uint4*>(W0 + ((int64_t)kc * d.dx + n) * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

struct StepTC {
  const float* P;
  const __nv_bfloat16* Wp;
  const float* obs;        // [R][n_obs]
  const float* c_in;       // [2A][R][64]
  const float* h_in;
  float* c_out;
  float* h_out;
  float* pi;               // [R][A][max_na]
  float* val;              // [R][A]
  int32_t* act;            // [R][A] or null
  float* zdbg;             // [2A][R][256] raw accumulators (debug) or null
  int64_t R;
  int done;                // pre-decision done flag
  int swap_lbo_sbo;        // debug: exchange the two descriptor strides
  uint32_t seed_lo, seed_hi, step;
  int64_t replica0;
  // optional activation store for the update (v2 only): bf16 [R/rc][2A][T][rc][w], w = dx / 256 / 64 / 64
  // (replica-chunk major, so that each update chunk is one contiguous block)
  __nv_bfloat16 *st_x, *st_g, *st_c, *st_h;
  int t, T;
  int64_t rc;
  // replica-range launches (v2 only): the per-unit state arrays have `ld` rows per unit (0: R) and the activation
  // store is indexed by the absolute replica row0 + r; every pointer is the base of the range's slice
  int64_t ld, row0;
  unsigned long long* prof;   // optional: per-phase clock64 sums of thread 0 of every CTA (tools/profile only)
};

extern __shared__ __align__(1024) unsigned char tc_smem[];

__global__ void __launch_bounds__(TC_THREADS, 1)
policy_step_tc_kernel(const DDimsTC d, const StepTC a) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = d.dx + TC_H, KC = K / 8, KS = K / 16;
  // ---- shared memory carve-up ----
  unsigned char* sB = tc_smem;                                  // KC * 4096
  unsigned char* sA = sB + (size_t)KC * 4096;                   // KC * 2048
  float* sStage = reinterpret_cast<float*>(sA + (size_t)KC * 2048);   // 16 x 64 floats; aliased by sRed [128][8]
  float* sWo = sStage + TC_STAGE_ROWS * TC_KTOT;                // [64][8]
  float* sBo = sWo + TC_H * 8;                                  // [8]
  float* sBias = sBo + 8;                                       // [256]
  uint64_t* sBar = reinterpret_cast<uint64_t*>(sBias + TC_N);   // mbarrier
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 1);      // TMEM base address
  float* sRed = sStage;

  const uint32_t bar = smem_u32(sBar);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *sTmem;
  // instruction descriptor: D=f32, A=B=bf16, K-major both, N=256, M=128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

  const int64_t n_tiles = (a.R + TC_M - 1) / TC_M;
  const int64_t n_items = n_tiles * 2 * d.A;
  const int64_t it_lo = n_items * blockIdx.x / gridDim.x, it_hi = n_items * (blockIdx.x + 1) / gridDim.x;
  int cur_u = -1;
  uint32_t parity = 0;
  // per-thread fc role: output column `col` of X
  const int col = tid;
  float w[TC_KW];
  float bias = 0.f;
  int kbase = 0, nk4 = 0, nw = 0, nt = 0, nf = 0, ooff = 0, n_in = 0, na = 0;

  for (int64_t it = it_lo; it < it_hi; ++it) {
    const int u = (int)(it / n_tiles);
    const int64_t r0 = (it - (int64_t)u * n_tiles) * TC_M;
    const int ag = u >> 1;
    if (u != cur_u) {
      cur_u = u;
      __syncthreads();
      // B operand image of this unit: plain 16-byte copies (async proxy will read it: fence below)
      const uint4* src = reinterpret_cast<const uint4*>(a.Wp + (int64_t)u * wp_stride(d.dx));
      uint4* dst = reinterpret_cast<uint4*>(sB);
      for (int i = tid; i < KC * TC_N; i += TC_THREADS) dst[i] = src[i];
      nw = d.n_wave[ag]; nt = d.n_wait[ag]; nf = d.ff > 0 ? d.n_fp[ag] : 0;
      n_in = nw + nt + nf; ooff = d.obs_off[ag]; na = d.n_a[ag];
#pragma unroll
      for (int k = 0; k < TC_KW; ++k) w[k] = 0.f;
      int nk = 0;
      if (col < d.dx) {
        if (col < d.fw) {
          nk = nw; kbase = 0;
#pragma unroll
          for (int k = 0; k < TC_KW; ++k) if (k < nw) w[k] = a.P[d.off_fcw_w[u] + (int64_t)k * d.fw + col];
          bias = a.P[d.off_fcw_b[u] + col];
        } else if (col < d.fw + d.ff) {
          nk = nf; kbase = TC_KW;
#pragma unroll
          for (int k = 0; k < TC_KF; ++k) if (k < nf) w[k] = a.P[d.off_fcf_w[u] + (int64_t)k * d.ff + (col - d.fw)];
          bias = a.P[d.off_fcf_b[u] + (col - d.fw)];
        } else {
          nk = nt; kbase = TC_KW + TC_KF;
#pragma unroll
          for (int k = 0; k < TC_KT; ++k) if (k < nt) w[k] = a.P[d.off_fct_w[u] + (int64_t)k * d.ft + (col - d.fw - d.ff)];
          bias = a.P[d.off_fct_b[u] + (col - d.fw - d.ff)];
        }
      }
      nk4 = (nk + 3) >> 2;
      for (int i = tid; i < TC_H * 8; i += TC_THREADS) {
        const int k = i >> 3, j = i & 7;
        sWo[i] = j < d.max_na ? a.P[d.off_wo + ((int64_t)u * TC_H + k) * d.max_na + j] : 0.f;
      }
      if (tid < 8) sBo[tid] = tid < d.max_na ? a.P[d.off_bo + (int64_t)u * d.max_na + tid] : 0.f;
      for (int i = tid; i < TC_N; i += TC_THREADS) sBias[i] = a.P[d.off_bl + (int64_t)u * TC_N + i];
    }
    // ---- 1. A tile: fc front end (cols 0..dx) in 16-row sub-blocks, then h_prev (cols dx..dx+64) ----
    for (int sb = 0; sb < TC_M / TC_STAGE_ROWS; ++sb) {
      __syncthreads();
      for (int i = tid; i < TC_STAGE_ROWS * TC_KTOT; i += TC_THREADS) sStage[i] = 0.f;
      __syncthreads();
      for (int i = tid; i < TC_STAGE_ROWS * n_in; i += TC_THREADS) {
        const int row = i / n_in, k = i - row * n_in;
        const int64_t r = r0 + sb * TC_STAGE_ROWS + row;
        if (r < a.R) {
          int dst;
          if (k < nw) dst = k;
          else if (k < nw + nt) dst = TC_KW + TC_KF + (k - nw);
          else dst = TC_KW + (k - nw - nt);
          sStage[row * TC_KTOT + dst] = a.obs[r * d.n_obs + ooff + k];
        }
      }
      __syncthreads();
      if (col < d.dx) {
        __nv_bfloat16* acol = reinterpret_cast<__nv_bfloat16*>(sA + (size_t)(col >> 3) * 2048) + (col & 7);
#pragma unroll 4
        for (int row = 0; row < TC_STAGE_ROWS; ++row) {
          const float4* in4 = reinterpret_cast<const float4*>(&sStage[row * TC_KTOT + kbase]);
          float acc = bias;
#pragma unroll
          for (int k4 = 0; k4 < TC_KW / 4; ++k4) {
            if (k4 < nk4) {
              const float4 x = in4[k4];
              acc = fmaf(x.x, w[4 * k4], acc); acc = fmaf(x.y, w[4 * k4 + 1], acc);
              acc = fmaf(x.z, w[4 * k4 + 2], acc); acc = fmaf(x.w, w[4 * k4 + 3], acc);
            }
          }
          acol[(sb * TC_STAGE_ROWS + row) * 8] = __float2bfloat16_rn(fmaxf(acc, 0.f));
        }
      }
    }
    {   // h_prev -> A columns dx .. dx+63 : thread = (row, 32-column half)
      const int row = tid >> 1, half = tid & 1;
      const int64_t r = r0 + row;
      const bool live = r < a.R && !a.done;
      const float4* hp = reinterpret_cast<const float4*>(a.h_in + ((int64_t)u * a.R + (r < a.R ? r : 0)) * TC_H + half * 32);
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        __align__(16) __nv_bfloat16 v[8];
        float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
        if (live) { x0 = hp[2 * c8]; x1 = hp[2 * c8 + 1]; }
        v[0] = __float2bfloat16_rn(x0.x); v[1] = __float2bfloat16_rn(x0.y); v[2] = __float2bfloat16_rn(x0.z); v[3] = __float2bfloat16_rn(x0.w);
        v[4] = __float2bfloat16_rn(x1.x); v[5] = __float2bfloat16_rn(x1.y); v[6] = __float2bfloat16_rn(x1.z); v[7] = __float2bfloat16_rn(x1.w);
        const int kc = (d.dx >> 3) + half * 4 + c8;
        *reinterpret_cast<uint4*>(sA + (size_t)kc * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v);
      }
    }
    // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- 2. MMA: D[128 x 256] = A[128 x K] . B[K x 256] ----
    if (warp == 0) {
      if (lane == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
        for (int ks = 0; ks < KS; ++ks) {
          uint64_t da, db;
          if (!a.swap_lbo_sbo) {
            da = make_desc(aA + ks * 2 * 2048, 2048, 128);
            db = make_desc(aB + ks * 2 * 4096, 4096, 128);
          } else {
            da = make_desc(aA + ks * 2 * 2048, 128, 2048);
            db = make_desc(aB + ks * 2 * 4096, 128, 4096);
          }
          umma_bf16(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- 3. epilogue: thread = (row = TMEM lane, 32 hidden units) ----
    {
      const int q = warp & 3, half = warp >> 2;
      const int row = q * 32 + lane;
      const int64_t r = r0 + row;
      const bool valid = r < a.R;
      const int64_t srow = ((int64_t)u * a.R + (valid ? r : 0)) * TC_H + half * 32;
      float lg[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) lg[j] = 0.f;
#pragma unroll
      for (int jb = 0; jb < 2; ++jb) {
        float zi[16], zf[16], zo[16], zu[16];
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32 + jb * 16);
        tmem_ld16(tbase, zi); tmem_ld16(tbase + 64, zf); tmem_ld16(tbase + 128, zo); tmem_ld16(tbase + 192, zu);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (a.zdbg && valid) {
          float* z = a.zdbg + ((int64_t)u * a.R + r) * TC_N + half * 32 + jb * 16;
#pragma unroll
          for (int e = 0; e < 16; ++e) { z[e] = zi[e]; z[64 + e] = zf[e]; z[128 + e] = zo[e]; z[192 + e] = zu[e]; }
        }
        float cprev[16];
        if (valid && !a.done) {
          const float4* cp = reinterpret_cast<const float4*>(a.c_in + srow + jb * 16);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 x = cp[e4];
            cprev[4 * e4] = x.x; cprev[4 * e4 + 1] = x.y; cprev[4 * e4 + 2] = x.z; cprev[4 * e4 + 3] = x.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) cprev[e] = 0.f;
        }
        float cn[16], hn[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int j = half * 32 + jb * 16 + e;
          const float gi = sigm(zi[e] + sBias[j]), gf = sigm(zf[e] + sBias[64 + j]);
          const float go = sigm(zo[e] + sBias[128 + j]), gu = tanh_fast(zu[e] + sBias[192 + j]);
          cn[e] = gf * cprev[e] + gi * gu;
          hn[e] = go * tanh_fast(cn[e]);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) lg[jj] = fmaf(hn[e], sWo[j * 8 + jj], lg[jj]);
        }
        if (valid) {
          float4* co = reinterpret_cast<float4*>(a.c_out + srow + jb * 16);
          float4* ho = reinterpret_cast<float4*>(a.h_out + srow + jb * 16);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            co[e4] = make_float4(cn[4 * e4], cn[4 * e4 + 1], cn[4 * e4 + 2], cn[4 * e4 + 3]);
            ho[e4] = make_float4(hn[4 * e4], hn[4 * e4 + 1], hn[4 * e4 + 2], hn[4 * e4 + 3]);
          }
        }
      }
      // all TMEM reads of this tile are done before the next tile's MMA may overwrite the accumulator
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();     // also: sStage (aliased by sRed) is free
      if (half == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sRed[row * 8 + j] = lg[j];
      }
      __syncthreads();
      if (half == 0 && valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) lg[j] += sRed[row * 8 + j] + sBo[j];
        if ((u & 1) == 0) {        // policy unit
          float mx = -1e30f;
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < na) mx = fmaxf(mx, lg[j]);
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) { lg[j] = j < na ? __expf(lg[j] - mx) : 0.f; s += lg[j]; }
          const float inv = 1.0f / s;
          float* po = a.pi + ((int64_t)r * d.A + ag) * d.max_na;
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < d.max_na) po[j] = lg[j] * inv;
          if (a.act) {
            uint32_t hsh = pmix32(a.seed_lo ^ (a.step * 0x9E3779B1U));
            hsh = pmix32(hsh ^ a.seed_hi ^ ((uint32_t)(a.replica0 + r) * 0x85EBCA77U));
            hsh = pmix32(hsh ^ ((uint32_t)ag * 0xC2B2AE3DU));
            const float uu = (float)(hsh >> 8) * (1.0f / 16777216.0f);
            float cum = 0.f;
            int pick = na - 1;
            bool found = false;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < na) { cum += lg[j] * inv; if (!found && uu < cum) { pick = j; found = true; } }
            a.act[(int64_t)r * d.A + ag] = pick;
          }
        } else {
          a.val[(int64_t)r * d.A + ag] = lg[0];
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ===================================================================================================
static size_t tc_smem_bytes(int K) {
  const int KC = K / 8;
  return (size_t)KC * 4096 + (size_t)KC * 2048 + (TC_STAGE_ROWS * TC_KTOT + TC_H * 8 + 8 + TC_N) * 4 + 16;
}

extern "C" int tscl_pack_weights(tscl_handle* h, const float* params, void* wpack_bf16, void* stream) {
  if (!h || !params || !wpack_bf16) return tsc_set_error("tscl_pack_weights: bad argument");
  PCK(cudaSetDevice(tscl_device_of(h)));
  const DDimsTC& d = *tscl_dims_of(h);
  if ((d.dx % 16) != 0) return tsc_set_error("tscl_pack_weights: dx must be a multiple of 16");
  dim3 grid(8, 2 * d.A);
  pack_wxh_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, params, (__nv_bfloat16*)wpack_bf16);
  pack_fc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, params, (__nv_bfloat16*)wpack_bf16);
  PCK(cudaGetLastError());
  return 0;
}

extern "C" int tscl_policy_step(tscl_handle* h, const float* params, const void* wpack_bf16, const float* obs,
                                int64_t R, const float* c_in, const float* h_in, float* c_out, float* h_out, float* pi,
                                float* val, int32_t* act, int32_t done, uint64_t seed, int64_t step, int64_t replica0,
                                float* zdbg, int32_t swap_lbo_sbo, void* stream) {
  if (!h || !params || !wpack_bf16 || !obs || R <= 0) return tsc_set_error("tscl_policy_step: bad argument");
  PCK(cudaSetDevice(tscl_device_of(h)));
  const DDimsTC& d = *tscl_dims_of(h);
  const int K = d.dx + TC_H;
  if ((d.dx % 16) != 0 || d.dx > TC_THREADS) return tsc_set_error("tscl_policy_step: unsupported dx");
  if (d.kw != 32) return tsc_set_error("tscl_policy_step: v1 kernel supports wave widths <= 32 only (use tscl_policy_step_v2)");
  const size_t smem = tc_smem_bytes(K);
  if (smem > 232448) return tsc_set_error("tscl_policy_step: operand tiles exceed shared memory");
  static int attr_dev = -1;
  if (attr_dev != tscl_device_of(h)) {
    PCK(cudaFuncSetAttribute(policy_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev = tscl_device_of(h);
  }
  int n_sm = 0;
  PCK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, tscl_device_of(h)));
  const int64_t n_items = ((R + TC_M - 1) / TC_M) * 2 * d.A;
  const int grid = (int)(n_items < n_sm ? n_items : n_sm);
  StepTC a;
  a.P = params; a.Wp = (const __nv_bfloat16*)wpack_bf16; a.obs = obs; a.c_in = c_in; a.h_in = h_in; a.c_out = c_out;
  a.h_out = h_out; a.pi = pi; a.val = val; a.act = act; a.zdbg = zdbg; a.R = R; a.done = done;
  a.swap_lbo_sbo = swap_lbo_sbo; a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  a.step = (uint32_t)step; a.replica0 = replica0;
  a.st_x = a.st_g = a.st_c = a.st_h = nullptr; a.t = 0; a.T = 1; a.rc = R; a.prof = nullptr;
  policy_step_tc_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(d, a);
  PCK(cudaGetLastError());
  return 0;
}

// ===================================================================================================
// v2: the fc front end also runs on the tensor cores.
//   A0 = observation slice tile [128 x 64] (wave32|fp16|wait16, bf16) staged in the LAST 8 K-chunks of the A tile,
//   B0 = block-diagonal fc weights [64 x dx] staged in the FIRST chunks of the A tile (both regions are dead
//        until the fc result is written / the h part is loaded), D0 = TMEM columns 256..256+dx;
//   TMEM -> registers -> (+bias, relu, bf16) -> A tile chunks 0..dx/8, then h_prev -> last 8 chunks, then the
//   gate MMA and the epilogue exactly as in v1.  Two tcgen05.commit per tile on one mbarrier.
// NT = 256: thread = (replica row, 32 hidden units); NT = 512: thread = (replica row, 16 hidden units) — twice the warps
// per SM for the latency-bound staging / epilogue phases (one CTA per SM either way: the weight operand fills shared
// memory), same arithmetic per element, so both variants produce identical bits.
// A-tile region of the v2 kernel: the operand tile, or the largest copy-out staging pass (128 rows x 528 B + the parked
// head partial sums behind the X pass), whichever is larger
__host__ __device__ inline size_t tc2_a_bytes(int K) {
  const size_t a = (size_t)(K / 8) * 2048;
  return a > 69632 ? a : 69632;
}
template <int NT, bool PROF>
__global__ void __launch_bounds__(NT, 1)
policy_step_tc2_kernel(const DDimsTC d, const StepTC a) {
  constexpr int NG = NT / 128;                 // hidden-unit groups per row (2 or 4)
  constexpr int HPT = TC_H / NG;               // hidden units per thread (32 or 16)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = d.dx + TC_H, KC = K / 8, KS = K / 16, KCX = d.dx / 8;
  unsigned char* sB = tc_smem;                                  // KC * 4096
  unsigned char* sA = sB + (size_t)KC * 4096;                   // KC * 2048
  float* sRed = reinterpret_cast<float*>(sA + tc2_a_bytes(K));      // [128][8]
  float* sWo = sRed + TC_M * 8;                                 // [64][8]
  float* sBo = sWo + TC_H * 8;                                  // [8]
  float* sBias = sBo + 8;                                       // [256]  lstm bias
  float* sBias0 = sBias + TC_N;                                 // [256]  fc biases
  uint64_t* sBar = reinterpret_cast<uint64_t*>(sBias0 + TC_N);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 2);
  const uint32_t bar = smem_u32(sBar), bar_fc = smem_u32(sBar + 1);   // MMA completion; fc-weight bulk copy landed
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar, 1); mbar_init(bar_fc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *sTmem;
  const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d.dx >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  const int64_t n_tiles = (a.R + TC_M - 1) / TC_M;
  const int64_t ld = a.ld > 0 ? a.ld : a.R;
  const int64_t n_items = n_tiles * 2 * d.A;
  const int64_t it_lo = n_items * blockIdx.x / gridDim.x, it_hi = n_items * (blockIdx.x + 1) / gridDim.x;
  int cur_u = -1;
  uint32_t parity = 0, par_fc = 0;
  int nw = 0, nt = 0, nf = 0, ooff = 0, na = 0, src_a = -1, src_b = -1;
  const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);

  long long pt[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pc = 0;
#define PROF_MARK(i) do { if (PROF && tid == 0) { const long long c_ = clock64(); pt[i] += c_ - pc; pc = c_; } } while (0)
  if (PROF && tid == 0) pc = clock64();
  for (int64_t it = it_lo; it < it_hi; ++it) {
    const int u = (int)(it / n_tiles);
    const int64_t r0 = (it - (int64_t)u * n_tiles) * TC_M;
    const int ag = u >> 1;
    // row of the activation store (chunk-outermost [R/rc][2A][T][rc][w]) for tile row `row`: one division per item
    const int64_t st_c0 = (a.row0 + r0) / a.rc, st_rin0 = (a.row0 + r0) - st_c0 * a.rc;
    auto store_row = [&](int row) -> int64_t {
      int64_t c = st_c0, rin = st_rin0 + row;
      while (rin >= a.rc) { rin -= a.rc; ++c; }
      return ((c * (2 * d.A) + u) * a.T + a.t) * a.rc + rin;
    };
    const __nv_bfloat16* Wu = a.Wp + (int64_t)u * wp_stride(d.dx);
    // NT = 512: no block-wide barrier here.  Warps 0-3 may still be in the previous tile's head softmax / sampling (they read
    // sRed, sRed2 = sA + 32 KB, sBo) while warps 4-15 already fetch and stage this tile's observation slice (last 8 chunks
    // of sA) and the fc-weight block (first 28 KB of sA); the barrier after the staging retires the previous tile.
    if (NT != 512 || u != cur_u) __syncthreads();      // previous tile fully retired (sRed, sA, sBo, TMEM readers)
    if (u != cur_u) {
      cur_u = u;
      const uint4* src = reinterpret_cast<const uint4*>(Wu);
      uint4* dst = reinterpret_cast<uint4*>(sB);
      for (int i = tid; i < KC * TC_N; i += NT) dst[i] = src[i];
      nw = d.n_wave[ag]; nt = d.n_wait[ag]; nf = d.ff > 0 ? d.n_fp[ag] : 0;
      ooff = d.obs_off[ag]; na = d.n_a[ag];
      {      // observation index of this lane's two input slots (-1: unused slot)
        auto slot_src = [&](int c) -> int {
          if (c < d.kw) return c < nw ? c : -1;
          if (c < d.kw + TC_KF) return c - d.kw < nf ? nw + nt + (c - d.kw) : -1;
          return c - d.kw - TC_KF < nt ? nw + (c - d.kw - TC_KF) : -1;
        };
        src_a = slot_src(2 * lane); src_b = slot_src(2 * lane + 1);
      }
      for (int i = tid; i < TC_H * 8; i += NT) {
        const int k = i >> 3, j = i & 7;
        sWo[i] = j < d.max_na ? a.P[d.off_wo + ((int64_t)u * TC_H + k) * d.max_na + j] : 0.f;
      }
      if (tid < 8) sBo[tid] = tid < d.max_na ? a.P[d.off_bo + (int64_t)u * d.max_na + tid] : 0.f;
      for (int i = tid; i < TC_N; i += NT) {
        sBias[i] = a.P[d.off_bl + (int64_t)u * TC_N + i];
        float b0 = 0.f;
        if (i < d.fw) b0 = a.P[d.off_fcw_b[u] + i];
        else if (i < d.fw + d.ff) b0 = a.P[d.off_fcf_b[u] + (i - d.fw)];
        else if (i < d.dx) b0 = a.P[d.off_fct_b[u] + (i - d.fw - d.ff)];
        sBias0[i] = b0;
      }
    }
    PROF_MARK(0);      // item-top sync + per-unit weight / constant (re)load
    // L2 prefetch of the NEXT work item's operands (observation slice, c, h): they are consumed ~one tile later
    if (it + 1 < it_hi) {
      const int un = (int)((it + 1) / n_tiles);
      const int64_t rn = ((it + 1) - (int64_t)un * n_tiles) * TC_M + (tid / NG);
      if (rn < a.R && (tid % NG) < 2) {
        const int an = un >> 1;
        const float* op = a.obs + rn * d.n_obs + d.obs_off[an] + (tid % NG) * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(op));
        const int64_t so = ((int64_t)un * ld + rn) * TC_H + (tid % NG) * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.c_in + so));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.h_in + so));
      }
    }
    // ---- 1a. B0 (fc weights) -> first chunks of the A tile; 1b. observation slice -> last 8 chunks ----
    {
      // fc weights (one contiguous 8 * dx * 16-byte block of the packed image) by ONE bulk copy (TMA unit, completion on
      // bar_fc): it overlaps the observation staging below; only the MMA-issuing thread waits for it
      // NT = 512: the copy for every tile but the CTA's first was issued right after the previous tile's gate MMA had
      // completed (the whole epilogue hides its ~3 k cycles)
      if (tid == NT - 1 && (NT != 512 || it == it_lo)) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy stores to the A tile precede this async write
        mbar_expect_tx(bar_fc, (uint32_t)(8 * d.dx * 16));
        bulk_g2s(aA, reinterpret_cast<const unsigned char*>(Wu + (int64_t)KC * TC_N * 8), (uint32_t)(8 * d.dx * 16), bar_fc);
      }
      if constexpr (NT == 512) {
        // a warp takes 8 rows; lane l owns input slots 2l, 2l + 1 of every row (their observation indices src_a / src_b were
        // resolved when the unit changed).  The slice of a row is one contiguous <= 256 B run of the observation vector, so
        // a warp load touches 2-3 sectors of ONE row (the former thread = (row, chunk) mapping touched 32 rows per
        // instruction: 8 k sector requests per tile, the whole staging phase); the two values leave as one packed store.
        // warps 4-15 share the 128 rows (11 each); warps 0-3 finish the previous tile's heads meanwhile
        if (warp >= 4) {
          const int rb = (warp - 4) * 11;
          float xa[11], xb[11];
#pragma unroll
          for (int rr = 0; rr < 11; ++rr) {
            const int64_t r = r0 + rb + rr;
            const bool ok = rb + rr < TC_M && r < a.R;
            const float* op = a.obs + (ok ? r : 0) * d.n_obs + ooff;
            xa[rr] = (src_a >= 0 && ok) ? __ldg(op + src_a) : 0.f;
            xb[rr] = (src_b >= 0 && ok) ? __ldg(op + src_b) : 0.f;
          }
#pragma unroll
          for (int rr = 0; rr < 11; ++rr) {
            const int row = rb + rr;
            if (row < TC_M) {
              const __nv_bfloat162 v = __floats2bfloat162_rn(xa[rr], xb[rr]);
              *reinterpret_cast<__nv_bfloat162*>(sA + (size_t)(KC - 8 + (lane >> 2)) * 2048 + row * 16 + (lane & 3) * 4) = v;
            }
          }
        }
      } else {
        // thread = (row, 16-byte chunk) : 128 x 8 pairs, 4 per thread
#pragma unroll
        for (int p = 0; p < 1024 / NT; ++p) {
          const int pair = p * NT + tid;
          const int row = pair & 127, ch = pair >> 7;          // consecutive threads -> consecutive rows
          const int64_t r = r0 + row;
          __align__(16) __nv_bfloat16 v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = ch * 8 + e;
            int src = -1;
            if (c < d.kw) { if (c < nw) src = c; }
            else if (c < d.kw + TC_KF) { if (c - d.kw < nf) src = nw + nt + (c - d.kw); }
            else { if (c - d.kw - TC_KF < nt) src = nw + (c - d.kw - TC_KF); }
            float x = 0.f;
            if (src >= 0 && r < a.R) x = __ldg(a.obs + r * d.n_obs + ooff + src);
            v[e] = __float2bfloat16_rn(x);
          }
          *reinterpret_cast<uint4*>(sA + (size_t)(KC - 8 + ch) * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    PROF_MARK(1);      // staging: fc weights + observation slice
    // ---- 2. MMA0: D0[128 x dx] = A0[128 x 64] . B0[64 x dx]  (TMEM columns 256..) ----
    if (warp == 0) {
      if (lane == 0) {
        mbar_wait(bar_fc, par_fc);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t da = make_desc(aA + (KC - 8 + 2 * ks) * 2048, 2048, 128);
          const uint64_t db = make_desc(aA + ks * 2 * (d.dx * 16), d.dx * 16, 128);
          umma_bf16(tmem + 256, da, db, idesc0, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    par_fc ^= 1;
    // h_{t-1} of this thread's (row, hidden-unit group): requested before the MMA wait, staged after the relu epilogue
    float4 hpre[HPT / 4];
    {
      const int64_t r = r0 + tid / NG;
      const bool live = r < a.R && !a.done;
      const float4* hp = reinterpret_cast<const float4*>(a.h_in + ((int64_t)u * ld + (r < a.R ? r : 0)) * TC_H + (tid % NG) * HPT);
#pragma unroll
      for (int i = 0; i < HPT / 4; ++i) hpre[i] = live ? hp[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    PROF_MARK(2);      // MMA0 issue + wait
    // ---- 3. X = relu(D0 + b) as bf16 -> A tile chunks 0..dx/8 ; h_prev -> last 8 chunks ----
    {
      const int q = warp & 3, hw = warp >> 2;
      const int row = q * 32 + lane;
      const int ncol = d.dx / NG;                      // columns per group (multiple of 8: dx % 32 == 0)
      const bool st = a.st_x && r0 + row < a.R;
      const int64_t m = st ? store_row(row) : 0;
      const int cend = (hw + 1) * ncol;
      for (int c0 = hw * ncol; c0 < cend;) {
        if (NT == 512 && (c0 & 15) == 0 && c0 + 16 <= cend) {      // 16 columns: one 256-bit store of the activation row
          float z[16];
          tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)c0, z);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          __align__(32) __nv_bfloat16 v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __float2bfloat16_rn(fmaxf(z[e] + sBias0[c0 + e], 0.f));
          *reinterpret_cast<uint4*>(sA + (size_t)(c0 >> 3) * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v);
          *reinterpret_cast<uint4*>(sA + (size_t)((c0 >> 3) + 1) * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v + 8);
          if (st) {
            const uint4 x = reinterpret_cast<const uint4*>(v)[0], y = reinterpret_cast<const uint4*>(v)[1];
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(a.st_x + (m * d.dx + c0)), "r"(x.x), "r"(x.y),
                         "r"(x.z), "r"(x.w), "r"(y.x), "r"(y.y), "r"(y.z), "r"(y.w) : "memory");
          }
          c0 += 16;
        } else {
          float z[8];
          tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)c0, z);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          __align__(16) __nv_bfloat16 v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(fmaxf(z[e] + sBias0[c0 + e], 0.f));
          *reinterpret_cast<uint4*>(sA + (size_t)(c0 >> 3) * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v);
          if (st) *reinterpret_cast<uint4*>(a.st_x + (m * d.dx + c0)) = *reinterpret_cast<const uint4*>(v);
          c0 += 8;
        }
      }
    }
    {
      const int row = tid / NG, half = tid % NG;      // `half` = hidden-unit group of HPT units
#pragma unroll
      for (int c8 = 0; c8 < HPT / 8; ++c8) {
        __align__(16) __nv_bfloat16 v[8];
        const float4 x0 = hpre[2 * c8], x1 = hpre[2 * c8 + 1];
        v[0] = __float2bfloat16_rn(x0.x); v[1] = __float2bfloat16_rn(x0.y); v[2] = __float2bfloat16_rn(x0.z); v[3] = __float2bfloat16_rn(x0.w);
        v[4] = __float2bfloat16_rn(x1.x); v[5] = __float2bfloat16_rn(x1.y); v[6] = __float2bfloat16_rn(x1.z); v[7] = __float2bfloat16_rn(x1.w);
        *reinterpret_cast<uint4*>(sA + (size_t)(KCX + half * (HPT / 8) + c8) * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    PROF_MARK(3);      // relu epilogue of the fc GEMM (+ st_x) and h staging
    // ---- 4. MMA1: gates D1[128 x 256] = [X | h][128 x K] . [Wx;Wh][K x 256] ----
    if (warp == 0) {
      if (lane == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ks = 0; ks < KS; ++ks)
          umma_bf16(tmem, make_desc(aA + ks * 2 * 2048, 2048, 128), make_desc(aB + ks * 2 * 4096, 4096, 128), idesc1,
                    ks > 0 ? 1u : 0u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    float cpre[16];        // c_{t-1} of this thread's 16 hidden units (NT = 512): in flight while the gate MMA runs
    if constexpr (NT == 512) {
      const int64_t r = r0 + (warp & 3) * 32 + lane;
      if (r < a.R && !a.done) {
        const float* cp = a.c_in + ((int64_t)u * ld + r) * TC_H + (warp >> 2) * HPT;
#pragma unroll
        for (int e8 = 0; e8 < 2; ++e8) {      // 256-bit loads: one full sector per thread and instruction
          uint32_t w[8];
          asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                       : "l"(cp + 8 * e8));
#pragma unroll
          for (int e = 0; e < 8; ++e) cpre[8 * e8 + e] = __uint_as_float(w[e]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) cpre[e] = 0.f;
      }
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (NT == 512 && tid == NT - 1 && it + 1 < it_hi) {
      // the A tile is dead: fetch the NEXT tile's fc-weight block now (its unit may differ)
      const int un = (int)((it + 1) / n_tiles);
      const __nv_bfloat16* Wn = a.Wp + (int64_t)un * wp_stride(d.dx);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar_fc, (uint32_t)(8 * d.dx * 16));
      bulk_g2s(aA, reinterpret_cast<const unsigned char*>(Wn + (int64_t)KC * TC_N * 8), (uint32_t)(8 * d.dx * 16), bar_fc);
    }
    PROF_MARK(4);      // gate MMA issue + wait
    // ---- 5. epilogue (identical to v1) ----
    {
      const int q = warp & 3, half = warp >> 2;
      const int row = q * 32 + lane;
      const int64_t r = r0 + row;
      const bool valid = r < a.R;
      const int64_t srow = ((int64_t)u * ld + (valid ? r : 0)) * TC_H + half * HPT;
      float lg[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) lg[j] = 0.f;
      // partial head sums of groups 1..NG-1: group 1 in sRed, groups 2, 3 (NT = 512) in the A tile, which is dead once the
      // gate MMA has been committed; group 0 adds them in a fixed order (deterministic bits)
      float* sRed2 = reinterpret_cast<float*>(sA + 32768);      // behind the fc-weight block, before the observation chunks
      auto finish_heads = [&]() {
        if (NG == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) lg[j] += sRed[row * 8 + j] + sBo[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            lg[j] = ((lg[j] + sRed[row * 8 + j]) + (sRed2[row * 8 + j] + sRed2[(TC_M + row) * 8 + j])) + sBo[j];
        }
        if ((u & 1) == 0) {
          float mx = -1e30f;
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < na) mx = fmaxf(mx, lg[j]);
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) { lg[j] = j < na ? __expf(lg[j] - mx) : 0.f; s += lg[j]; }
          const float inv = 1.0f / s;
          float* po = a.pi + ((int64_t)r * d.A + ag) * d.max_na;
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < d.max_na) po[j] = lg[j] * inv;
          if (a.act) {
            uint32_t hsh = pmix32(a.seed_lo ^ (a.step * 0x9E3779B1U));
            hsh = pmix32(hsh ^ a.seed_hi ^ ((uint32_t)(a.replica0 + r) * 0x85EBCA77U));
            hsh = pmix32(hsh ^ ((uint32_t)ag * 0xC2B2AE3DU));
            const float uu = (float)(hsh >> 8) * (1.0f / 16777216.0f);
            float cum = 0.f;
            int pick = na - 1;
            bool found = false;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < na) { cum += lg[j] * inv; if (!found && uu < cum) { pick = j; found = true; } }
            a.act[(int64_t)r * d.A + ag] = pick;
          }
        } else {
          a.val[(int64_t)r * d.A + ag] = lg[0];
        }
      };
#pragma unroll
      for (int jb = 0; jb < HPT / 16; ++jb) {
        float zi[16], zf[16], zo[16], zu[16];
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HPT + jb * 16);
        tmem_ld16(tbase, zi); tmem_ld16(tbase + 64, zf); tmem_ld16(tbase + 128, zo); tmem_ld16(tbase + 192, zu);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (a.zdbg && valid) {
          float* z = a.zdbg + ((int64_t)u * ld + r) * TC_N + half * HPT + jb * 16;
#pragma unroll
          for (int e = 0; e < 16; ++e) { z[e] = zi[e]; z[64 + e] = zf[e]; z[128 + e] = zo[e]; z[192 + e] = zu[e]; }
        }
        float cprev[16];
        if constexpr (NT == 512) {
#pragma unroll
          for (int e = 0; e < 16; ++e) cprev[e] = cpre[e];
        } else if (valid && !a.done) {
          const float4* cp = reinterpret_cast<const float4*>(a.c_in + srow + jb * 16);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 x = cp[e4];
            cprev[4 * e4] = x.x; cprev[4 * e4 + 1] = x.y; cprev[4 * e4 + 2] = x.z; cprev[4 * e4 + 3] = x.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) cprev[e] = 0.f;
        }
        float cn[16], hn[16];
        __align__(16) __nv_bfloat16 gbuf[4][16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int j = half * HPT + jb * 16 + e;
          const float gi = sigm(zi[e] + sBias[j]), gf = sigm(zf[e] + sBias[64 + j]);
          const float go = sigm(zo[e] + sBias[128 + j]), gu = tanh_fast(zu[e] + sBias[192 + j]);
          cn[e] = gf * cprev[e] + gi * gu;
          hn[e] = go * tanh_fast(cn[e]);
          gbuf[0][e] = __float2bfloat16_rn(gi); gbuf[1][e] = __float2bfloat16_rn(gf);
          gbuf[2][e] = __float2bfloat16_rn(go); gbuf[3][e] = __float2bfloat16_rn(gu);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) lg[jj] = fmaf(hn[e], sWo[j * 8 + jj], lg[jj]);
        }
        if constexpr (NT == 512) {
          // Stores: a thread owns (row, 16 hidden units), i.e. 32-byte pieces of its row.  They leave as 256-bit stores
          // (STG.256: one full 32-byte sector per thread and instruction, 14 instructions per thread and tile) straight
          // from the registers, interleaved with the other warps' cell math.  Measured alternatives: 16-byte row-owner
          // stores (28 half-sector instructions: the LSU was ~half of the kernel), a SIMT copy-out through XOR-swizzled
          // shared memory (coalesced, but 4 staging passes with 8 block-wide barriers: 16 k cycles per tile) and the same
          // passes handed to the copy engine row by row (cp.async.bulk, 128-512 B each: slower still).
          auto st32 = [](void* p, const void* v) {
            const uint4 x = reinterpret_cast<const uint4*>(v)[0], y = reinterpret_cast<const uint4*>(v)[1];
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w),
                         "r"(y.x), "r"(y.y), "r"(y.z), "r"(y.w) : "memory");
          };
          if (valid) {
            if (a.st_g) {
              const int64_t m = store_row(row);
#pragma unroll
              for (int g = 0; g < 4; ++g) st32(a.st_g + m * TC_N + g * 64 + half * HPT, &gbuf[g][0]);
              __align__(32) __nv_bfloat16 cb[16], hb[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) { cb[e] = __float2bfloat16_rn(cn[e]); hb[e] = __float2bfloat16_rn(hn[e]); }
              st32(a.st_c + m * TC_H + half * HPT, cb);
              st32(a.st_h + m * TC_H + half * HPT, hb);
            }
            float* co = a.c_out + srow;
            float* ho = a.h_out + srow;
            st32(co, cn); st32(co + 8, cn + 8);
            st32(ho, hn); st32(ho + 8, hn + 8);
          }
        } else {
        if (valid && a.st_g) {
          const int64_t m = store_row((int)(r - r0));
          const int jo = half * HPT + jb * 16;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4* o = reinterpret_cast<uint4*>(a.st_g + m * TC_N + g * 64 + jo);
            o[0] = *reinterpret_cast<const uint4*>(&gbuf[g][0]); o[1] = *reinterpret_cast<const uint4*>(&gbuf[g][8]);
          }
          __align__(16) __nv_bfloat16 cb[16], hb[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) { cb[e] = __float2bfloat16_rn(cn[e]); hb[e] = __float2bfloat16_rn(hn[e]); }
          uint4* oc = reinterpret_cast<uint4*>(a.st_c + m * TC_H + jo);
          uint4* oh = reinterpret_cast<uint4*>(a.st_h + m * TC_H + jo);
          oc[0] = *reinterpret_cast<const uint4*>(cb); oc[1] = *reinterpret_cast<const uint4*>(cb + 8);
          oh[0] = *reinterpret_cast<const uint4*>(hb); oh[1] = *reinterpret_cast<const uint4*>(hb + 8);
        }
        if (valid) {
          float4* co = reinterpret_cast<float4*>(a.c_out + srow + jb * 16);
          float4* ho = reinterpret_cast<float4*>(a.h_out + srow + jb * 16);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            co[e4] = make_float4(cn[4 * e4], cn[4 * e4 + 1], cn[4 * e4 + 2], cn[4 * e4 + 3]);
            ho[e4] = make_float4(hn[4 * e4], hn[4 * e4 + 1], hn[4 * e4 + 2], hn[4 * e4 + 3]);
          }
        }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (half == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sRed[row * 8 + j] = lg[j];
      } else if (half > 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sRed2[((half - 2) * TC_M + row) * 8 + j] = lg[j];
      }
      __syncthreads();
      PROF_MARK(5);    // cell + stores + head partial sums
      if (half == 0 && valid) finish_heads();
    }
  }
  PROF_MARK(6);        // head softmax / sampling of the last item
  if (PROF && tid == 0)
    for (int i = 0; i < 16; ++i) atomicAdd(a.prof + i, (unsigned long long)pt[i]);
#undef PROF_MARK
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ===================================================================================================
// v3: the same fused step with the weight operand STREAMED instead of resident, so that TWO CTAs share an SM and
// overlap each other's serial phases (v2 holds the 147 KB [Wx;Wh] image in shared memory: one CTA per SM, every
// latency exposed).
//   * [Wx;Wh] of the current unit stays in L2 (7.4 MB for all 50 units); its 18 K-slabs of 8 KB travel through a
//     4-stage ring in shared memory by cp.async.bulk (TMA unit, mbarrier complete_tx), issued and consumed by ONE
//     thread, which also issues the tcgen05.mma of each slab and commits it to the slab's "empty" barrier;
//     the first slabs of the NEXT work item and its fc-weight block are prefetched while the current item is in its
//     epilogue;
//   * TMEM: 256 columns per CTA (two CTAs = 512): the fc accumulator D0 (dx columns) and the gate accumulator D1 (256)
//     share them, D0 is dead once X has been written to the A tile;
//   * shared memory 112 KB per CTA: A tile 72 KB | ring 32 KB | biases, head weights, partial head sums.
// Arithmetic per element is identical to v2: the two kernels produce the same bits.
#define V3_NSTG 4
#define V3_SLAB 8192
__global__ void __launch_bounds__(256, 2)
policy_step_tc3_kernel(const DDimsTC d, const StepTC a) {
  constexpr int NT = 256, NG = 2, HPT = 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = d.dx + TC_H, KC = K / 8, KS = K / 16, KCX = d.dx / 8;
  unsigned char* sA = tc_smem;                                  // KC * 2048
  unsigned char* sR = sA + (size_t)KC * 2048;                   // ring: V3_NSTG * 8192
  float* sRed = reinterpret_cast<float*>(sR + V3_NSTG * V3_SLAB);   // [128][8]
  float* sWo = sRed + TC_M * 8;                                 // [64][8]
  float* sBo = sWo + TC_H * 8;                                  // [8]
  float* sBias = sBo + 8;                                       // [256]  lstm bias
  float* sBias0 = sBias + TC_N;                                 // [256]  fc biases
  uint64_t* sBar = reinterpret_cast<uint64_t*>(sBias0 + TC_N);  // [0] mma done, [1] fc weights landed, [2..5] full, [6..9] empty
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 2 + 2 * V3_NSTG);
  const uint32_t bar_mma = smem_u32(sBar), bar_fc = smem_u32(sBar + 1);
  const uint32_t bar_full = smem_u32(sBar + 2), bar_empty = smem_u32(sBar + 2 + V3_NSTG);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar_mma, 1); mbar_init(bar_fc, 1);
    for (int s = 0; s < V3_NSTG; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *sTmem;
  const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d.dx >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  const int64_t n_tiles = (a.R + TC_M - 1) / TC_M;
  const int64_t ld = a.ld > 0 ? a.ld : a.R;
  const int64_t n_items = n_tiles * 2 * d.A;
  const int64_t it_lo = n_items * blockIdx.x / gridDim.x, it_hi = n_items * (blockIdx.x + 1) / gridDim.x;
  int cur_u = -1;
  uint32_t par_mma = 0, par_fc = 0;
  uint32_t n_issue = 0, n_cons = 0;            // thread 0: slabs issued / consumed since the kernel started
  int nw = 0, nt = 0, nf = 0, ooff = 0, na = 0;
  const uint32_t aA = smem_u32(sA), aR = smem_u32(sR);
  const uint32_t fc_bytes = (uint32_t)(8 * d.dx * 16);

  // thread 0 only: stream slab `ks` of unit `u` into the next ring stage (waits until the MMA that last read it is done)
  auto issue_slab = [&](int u, int ks) {
    const uint32_t s = n_issue % V3_NSTG, k = n_issue / V3_NSTG;
    if (k > 0) mbar_wait(bar_empty + 8 * s, (k - 1) & 1);
    mbar_expect_tx(bar_full + 8 * s, V3_SLAB);
    bulk_g2s(aR + s * V3_SLAB, reinterpret_cast<const unsigned char*>(a.Wp + (int64_t)u * wp_stride(d.dx)) + (size_t)ks * V3_SLAB,
             V3_SLAB, bar_full + 8 * s);
    ++n_issue;
  };
  const int n_pre = KS < V3_NSTG ? KS : V3_NSTG;
  if (tid == 0 && it_lo < it_hi) {             // first work item: fc weights + the first slabs
    const int u0 = (int)(it_lo / n_tiles);
    mbar_expect_tx(bar_fc, fc_bytes);
    bulk_g2s(aA, reinterpret_cast<const unsigned char*>(a.Wp + (int64_t)u0 * wp_stride(d.dx)) + (size_t)KC * 4096, fc_bytes, bar_fc);
    for (int ks = 0; ks < n_pre; ++ks) issue_slab(u0, ks);
  }

  for (int64_t it = it_lo; it < it_hi; ++it) {
    const int u = (int)(it / n_tiles);
    const int64_t r0 = (it - (int64_t)u * n_tiles) * TC_M;
    const int ag = u >> 1;
    const int64_t st_c0 = (a.row0 + r0) / a.rc, st_rin0 = (a.row0 + r0) - st_c0 * a.rc;
    auto store_row = [&](int row) -> int64_t {
      int64_t c = st_c0, rin = st_rin0 + row;
      while (rin >= a.rc) { rin -= a.rc; ++c; }
      return ((c * (2 * d.A) + u) * a.T + a.t) * a.rc + rin;
    };
    if (u != cur_u) {      // per-unit constants (sRed / sWo / biases are not touched by the async copies in flight)
      __syncthreads();     // previous item's epilogue has finished reading them
      cur_u = u;
      nw = d.n_wave[ag]; nt = d.n_wait[ag]; nf = d.ff > 0 ? d.n_fp[ag] : 0;
      ooff = d.obs_off[ag]; na = d.n_a[ag];
      for (int i = tid; i < TC_H * 8; i += NT) {
        const int k = i >> 3, j = i & 7;
        sWo[i] = j < d.max_na ? a.P[d.off_wo + ((int64_t)u * TC_H + k) * d.max_na + j] : 0.f;
      }
      if (tid < 8) sBo[tid] = tid < d.max_na ? a.P[d.off_bo + (int64_t)u * d.max_na + tid] : 0.f;
      for (int i = tid; i < TC_N; i += NT) {
        sBias[i] = a.P[d.off_bl + (int64_t)u * TC_N + i];
        float b0 = 0.f;
        if (i < d.fw) b0 = a.P[d.off_fcw_b[u] + i];
        else if (i < d.fw + d.ff) b0 = a.P[d.off_fcf_b[u] + (i - d.fw)];
        else if (i < d.dx) b0 = a.P[d.off_fct_b[u] + (i - d.fw - d.ff)];
        sBias0[i] = b0;
      }
    }
    if (it + 1 < it_hi) {      // L2 prefetch of the next item's observation slice and state rows
      const int un = (int)((it + 1) / n_tiles);
      const int64_t rn = ((it + 1) - (int64_t)un * n_tiles) * TC_M + (tid / NG);
      if (rn < a.R) {
        const int an = un >> 1;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.obs + rn * d.n_obs + d.obs_off[an] + (tid % NG) * 32));
        const int64_t so = ((int64_t)un * ld + rn) * TC_H + (tid % NG) * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.c_in + so));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.h_in + so));
      }
    }
    // ---- 1. observation slice -> last 8 chunks of the A tile (the fc weights arrive by bulk copy in chunks 0..) ----
#pragma unroll
    for (int p = 0; p < 1024 / NT; ++p) {
      const int pair = p * NT + tid;
      const int row = pair & 127, ch = pair >> 7;
      const int64_t r = r0 + row;
      __align__(16) __nv_bfloat16 v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = ch * 8 + e;
        int src = -1;
        if (c < d.kw) { if (c < nw) src = c; }
        else if (c < d.kw + TC_KF) { if (c - d.kw < nf) src = nw + nt + (c - d.kw); }
        else { if (c - d.kw - TC_KF < nt) src = nw + (c - d.kw - TC_KF); }
        float x = 0.f;
        if (src >= 0 && r < a.R) x = __ldg(a.obs + r * d.n_obs + ooff + src);
        v[e] = __float2bfloat16_rn(x);
      }
      *reinterpret_cast<uint4*>(sA + (size_t)(KC - 8 + ch) * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- 2. MMA0: D0[128 x dx] = A0[128 x 64] . B0[64 x dx]  (TMEM columns 0..dx) ----
    if (tid == 0) {
      mbar_wait(bar_fc, par_fc);               // fc weights of this item have landed in chunks 0..
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t da = make_desc(aA + (KC - 8 + 2 * ks) * 2048, 2048, 128);
        const uint64_t db = make_desc(aA + ks * 2 * (d.dx * 16), d.dx * 16, 128);
        umma_bf16(tmem, da, db, idesc0, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    par_fc ^= 1;
    mbar_wait(bar_mma, par_mma);
    par_mma ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- 3. X = relu(D0 + b) as bf16 -> A tile chunks 0..dx/8 ; h_prev -> the 8 chunks after them ----
    {
      const int q = warp & 3, hw = warp >> 2;
      const int row = q * 32 + lane;
      const int ncol = d.dx / NG;
      const bool st = a.st_x && r0 + row < a.R;
      const int64_t m = st ? store_row(row) : 0;
      for (int c0 = hw * ncol; c0 < (hw + 1) * ncol; c0 += 8) {
        float z[8];
        tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, z);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(fmaxf(z[e] + sBias0[c0 + e], 0.f));
        *reinterpret_cast<uint4*>(sA + (size_t)(c0 >> 3) * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v);
        if (st) *reinterpret_cast<uint4*>(a.st_x + (m * d.dx + c0)) = *reinterpret_cast<const uint4*>(v);
      }
    }
    {
      const int row = tid / NG, half = tid % NG;
      const int64_t r = r0 + row;
      const bool live = r < a.R && !a.done;
      const float4* hp = reinterpret_cast<const float4*>(a.h_in + ((int64_t)u * ld + (r < a.R ? r : 0)) * TC_H + half * HPT);
#pragma unroll
      for (int c8 = 0; c8 < HPT / 8; ++c8) {
        __align__(16) __nv_bfloat16 v[8];
        float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
        if (live) { x0 = hp[2 * c8]; x1 = hp[2 * c8 + 1]; }
        v[0] = __float2bfloat16_rn(x0.x); v[1] = __float2bfloat16_rn(x0.y); v[2] = __float2bfloat16_rn(x0.z); v[3] = __float2bfloat16_rn(x0.w);
        v[4] = __float2bfloat16_rn(x1.x); v[5] = __float2bfloat16_rn(x1.y); v[6] = __float2bfloat16_rn(x1.z); v[7] = __float2bfloat16_rn(x1.w);
        *reinterpret_cast<uint4*>(sA + (size_t)(KCX + half * (HPT / 8) + c8) * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- 4. MMA1: gates D1[128 x 256] = [X | h][128 x K] . [Wx;Wh][K x 256], B slabs through the ring ----
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t s = n_cons % V3_NSTG, k = n_cons / V3_NSTG;
        mbar_wait(bar_full + 8 * s, k & 1);
        umma_bf16(tmem, make_desc(aA + ks * 2 * 2048, 2048, 128), make_desc(aR + s * V3_SLAB, 4096, 128), idesc1,
                  ks > 0 ? 1u : 0u);
        umma_commit(bar_empty + 8 * s);
        ++n_cons;
        if (ks >= 1 && ks - 1 + V3_NSTG < KS) issue_slab(u, ks - 1 + V3_NSTG);   // refill the stage MMA(ks-1) has left
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, par_mma);
    par_mma ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0 && it + 1 < it_hi) {          // the A tile and the ring are free: start the next item's operands now
      const int un = (int)((it + 1) / n_tiles);
      mbar_expect_tx(bar_fc, fc_bytes);
      bulk_g2s(aA, reinterpret_cast<const unsigned char*>(a.Wp + (int64_t)un * wp_stride(d.dx)) + (size_t)KC * 4096, fc_bytes, bar_fc);
      for (int ks = 0; ks < n_pre; ++ks) issue_slab(un, ks);
    }
    // ---- 5. epilogue (as v2, NT = 256) ----
    {
      const int q = warp & 3, half = warp >> 2;
      const int row = q * 32 + lane;
      const int64_t r = r0 + row;
      const bool valid = r < a.R;
      const int64_t srow = ((int64_t)u * ld + (valid ? r : 0)) * TC_H + half * HPT;
      float lg[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) lg[j] = 0.f;
#pragma unroll 1
      for (int jb = 0; jb < HPT / 16; ++jb) {
        float zi[16], zf[16], zo[16], zu[16];
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HPT + jb * 16);
        tmem_ld16(tbase, zi); tmem_ld16(tbase + 64, zf); tmem_ld16(tbase + 128, zo); tmem_ld16(tbase + 192, zu);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float cprev[16];
        if (valid && !a.done) {
          const float4* cp = reinterpret_cast<const float4*>(a.c_in + srow + jb * 16);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 x = cp[e4];
            cprev[4 * e4] = x.x; cprev[4 * e4 + 1] = x.y; cprev[4 * e4 + 2] = x.z; cprev[4 * e4 + 3] = x.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) cprev[e] = 0.f;
        }
        float cn[16], hn[16];
        __align__(16) __nv_bfloat16 gbuf[4][16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int j = half * HPT + jb * 16 + e;
          const float gi = sigm(zi[e] + sBias[j]), gf = sigm(zf[e] + sBias[64 + j]);
          const float go = sigm(zo[e] + sBias[128 + j]), gu = tanh_fast(zu[e] + sBias[192 + j]);
          cn[e] = gf * cprev[e] + gi * gu;
          hn[e] = go * tanh_fast(cn[e]);
          gbuf[0][e] = __float2bfloat16_rn(gi); gbuf[1][e] = __float2bfloat16_rn(gf);
          gbuf[2][e] = __float2bfloat16_rn(go); gbuf[3][e] = __float2bfloat16_rn(gu);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) lg[jj] = fmaf(hn[e], sWo[j * 8 + jj], lg[jj]);
        }
        if (valid && a.st_g) {
          const int64_t m = store_row((int)(r - r0));
          const int jo = half * HPT + jb * 16;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4* o = reinterpret_cast<uint4*>(a.st_g + m * TC_N + g * 64 + jo);
            o[0] = *reinterpret_cast<const uint4*>(&gbuf[g][0]); o[1] = *reinterpret_cast<const uint4*>(&gbuf[g][8]);
          }
          __align__(16) __nv_bfloat16 cb[16], hb[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) { cb[e] = __float2bfloat16_rn(cn[e]); hb[e] = __float2bfloat16_rn(hn[e]); }
          uint4* oc = reinterpret_cast<uint4*>(a.st_c + m * TC_H + jo);
          uint4* oh = reinterpret_cast<uint4*>(a.st_h + m * TC_H + jo);
          oc[0] = *reinterpret_cast<const uint4*>(cb); oc[1] = *reinterpret_cast<const uint4*>(cb + 8);
          oh[0] = *reinterpret_cast<const uint4*>(hb); oh[1] = *reinterpret_cast<const uint4*>(hb + 8);
        }
        if (valid) {
          float4* co = reinterpret_cast<float4*>(a.c_out + srow + jb * 16);
          float4* ho = reinterpret_cast<float4*>(a.h_out + srow + jb * 16);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            co[e4] = make_float4(cn[4 * e4], cn[4 * e4 + 1], cn[4 * e4 + 2], cn[4 * e4 + 3]);
            ho[e4] = make_float4(hn[4 * e4], hn[4 * e4 + 1], hn[4 * e4 + 2], hn[4 * e4 + 3]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (half == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sRed[row * 8 + j] = lg[j];
      }
      __syncthreads();       // also: every thread is done with TMEM -> the next item's MMA0 may overwrite D
      if (half == 0 && valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) lg[j] += sRed[row * 8 + j] + sBo[j];
        if ((u & 1) == 0) {
          float mx = -1e30f;
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < na) mx = fmaxf(mx, lg[j]);
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) { lg[j] = j < na ? __expf(lg[j] - mx) : 0.f; s += lg[j]; }
          const float inv = 1.0f / s;
          float* po = a.pi + ((int64_t)r * d.A + ag) * d.max_na;
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < d.max_na) po[j] = lg[j] * inv;
          if (a.act) {
            uint32_t hsh = pmix32(a.seed_lo ^ (a.step * 0x9E3779B1U));
            hsh = pmix32(hsh ^ a.seed_hi ^ ((uint32_t)(a.replica0 + r) * 0x85EBCA77U));
            hsh = pmix32(hsh ^ ((uint32_t)ag * 0xC2B2AE3DU));
            const float uu = (float)(hsh >> 8) * (1.0f / 16777216.0f);
            float cum = 0.f;
            int pick = na - 1;
            bool found = false;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < na) { cum += lg[j] * inv; if (!found && uu < cum) { pick = j; found = true; } }
            a.act[(int64_t)r * d.A + ag] = pick;
          }
        } else {
          a.val[(int64_t)r * d.A + ag] = lg[0];
        }
      }
      __syncthreads();       // sRed is rewritten by the next item's epilogue only, but its head reads must finish first
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static size_t tc3_smem_bytes(int K) {
  const int KC = K / 8;
  return (size_t)KC * 2048 + V3_NSTG * V3_SLAB + (TC_M * 8 + TC_H * 8 + 8 + TC_N + TC_N) * 4 + (2 + 2 * V3_NSTG) * 8 + 16;
}

static size_t tc2_smem_bytes(int K) {
  const int KC = K / 8;
  return (size_t)KC * 4096 + tc2_a_bytes(K) + (TC_M * 8 + TC_H * 8 + 8 + TC_N + TC_N) * 4 + 32;
}

static unsigned long long* g_policy_prof = nullptr;
// tools only: device pointer to 8 uint64 counters that receive per-phase clock64 sums of the v2 kernel (NULL = off)
extern "C" int tscl_debug_policy_prof(void* counters_dev) { g_policy_prof = (unsigned long long*)counters_dev; return 0; }
static unsigned long long* g_bptt_prof = nullptr;
extern "C" int tscl_debug_bptt_prof(void* counters_dev) { g_bptt_prof = (unsigned long long*)counters_dev; return 0; }

extern "C" int tscl_policy_step_v2r(tscl_handle* h, const float* params, const void* wpack_bf16, const float* obs,
                                   int64_t R, const float* c_in, const float* h_in, float* c_out, float* h_out,
                                   float* pi, float* val, int32_t* act, int32_t done, uint64_t seed, int64_t step,
                                   int64_t replica0, float* zdbg, void* st_x, void* st_g, void* st_c, void* st_h,
                                   int32_t t, int32_t T, int64_t rc, int64_t ld_state, int64_t row0, void* stream) {
  if (!h || !params || !wpack_bf16 || !obs || R <= 0) return tsc_set_error("tscl_policy_step_v2r: bad argument");
  if (st_x && (rc <= 0 || (ld_state > 0 ? ld_state : R) % rc != 0)) return tsc_set_error("tscl_policy_step_v2r: store chunk must divide the replica count");
  if (ld_state > 0 && (row0 < 0 || row0 + R > ld_state)) return tsc_set_error("tscl_policy_step_v2r: replica range outside [0, ld_state)");
  PCK(cudaSetDevice(tscl_device_of(h)));
  const DDimsTC& d = *tscl_dims_of(h);
  const int K = d.dx + TC_H;
  if ((d.dx % 32) != 0 || d.dx > 256) return tsc_set_error("tscl_policy_step_v2r: dx must be a multiple of 32, <= 256");
  if (d.kw == 0) return tsc_set_error("tscl_policy_step_v2r: observation slice does not fit the 64-column input tile");
  if (8 * d.dx * 16 > (d.dx / 8) * 2048) return tsc_set_error("tscl_policy_step_v2r: fc operand does not fit its staging region");
  const bool wide_tile = d.dx <= 224;   // NT = 512: the A-tile region doubles as the copy-out staging (X pitch dx * 2 + 16 <= 464 B)
  const size_t smem = tc2_smem_bytes(K);
  if (smem > 232448) return tsc_set_error("tscl_policy_step_v2r: operand tiles exceed shared memory");
  static int attr_dev = -1;
  if (attr_dev != tscl_device_of(h)) {
    PCK(cudaFuncSetAttribute(policy_step_tc2_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PCK(cudaFuncSetAttribute(policy_step_tc2_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PCK(cudaFuncSetAttribute(policy_step_tc2_kernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev = tscl_device_of(h);
  }
  int n_sm = 0;
  PCK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, tscl_device_of(h)));
  const int64_t n_items = ((R + TC_M - 1) / TC_M) * 2 * d.A;
  const int grid = (int)(n_items < n_sm ? n_items : n_sm);
  StepTC a;
  a.P = params; a.Wp = (const __nv_bfloat16*)wpack_bf16; a.obs = obs; a.c_in = c_in; a.h_in = h_in; a.c_out = c_out;
  a.h_out = h_out; a.pi = pi; a.val = val; a.act = act; a.zdbg = zdbg; a.R = R; a.done = done; a.swap_lbo_sbo = 0;
  a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32); a.step = (uint32_t)step; a.replica0 = replica0;
  a.st_x = (__nv_bfloat16*)st_x; a.st_g = (__nv_bfloat16*)st_g; a.st_c = (__nv_bfloat16*)st_c; a.st_h = (__nv_bfloat16*)st_h;
  a.t = t; a.T = T > 0 ? T : 1; a.rc = rc > 0 ? rc : R; a.ld = ld_state; a.row0 = ld_state > 0 ? row0 : 0;
  a.prof = g_policy_prof;
  // default: the resident-weight kernel v2 (512 threads; TSC_POLICY_THREADS=256 = its round-1 thread mapping).
  // TSC_POLICY_V3=1 selects the streamed-weight two-CTA-per-SM kernel: measured 1.404 vs 1.342 ms of rollout per control
  // step at R = 8192 (policy kernel ~0.63 vs 0.57 ms), i.e. slower, so it is not the default
  static const int pol_v3 = []() { const char* e = getenv("TSC_POLICY_V3"); return e ? atoi(e) : 0; }();
  static const int pol_threads = []() { const char* e = getenv("TSC_POLICY_THREADS"); return e && atoi(e) == 256 ? 256 : 512; }();
  if (pol_v3 && !zdbg) {
    const size_t smem3 = tc3_smem_bytes(K);
    static int attr3_dev = -1;
    if (attr3_dev != tscl_device_of(h)) {
      PCK(cudaFuncSetAttribute(policy_step_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
      attr3_dev = tscl_device_of(h);
    }
    const int grid3 = (int)(n_items < 2 * n_sm ? n_items : 2 * n_sm);       // two CTAs per SM
    policy_step_tc3_kernel<<<grid3, 256, smem3, (cudaStream_t)stream>>>(d, a);
  } else if (a.prof && wide_tile) policy_step_tc2_kernel<512, true><<<grid, 512, smem, (cudaStream_t)stream>>>(d, a);
  else if (pol_threads == 512 && wide_tile) policy_step_tc2_kernel<512, false><<<grid, 512, smem, (cudaStream_t)stream>>>(d, a);
  else policy_step_tc2_kernel<256, false><<<grid, 256, smem, (cudaStream_t)stream>>>(d, a);
  PCK(cudaGetLastError());
  return 0;
}

extern "C" int tscl_policy_step_v2(tscl_handle* h, const float* params, const void* wpack_bf16, const float* obs, int64_t R,
                                   const float* c_in, const float* h_in, float* c_out, float* h_out, float* pi, float* val,
                                   int32_t* act, int32_t done, uint64_t seed, int64_t step, int64_t replica0, float* zdbg,
                                   void* st_x, void* st_g, void* st_c, void* st_h, int32_t t, int32_t T, int64_t rc,
                                   void* stream) {
  return tscl_policy_step_v2r(h, params, wpack_bf16, obs, R, c_in, h_in, c_out, h_out, pi, val, act, done, seed, step, replica0,
                              zdbg, st_x, st_g, st_c, st_h, t, T, rc, 0, 0, stream);
}

// ===================================================================================================
// BPTT through the LSTM on the tensor cores (replaces lstm_seq_bwd_kernel of tsc_learn.cu).
// One CTA walks (unit, 128-replica tile) items; for t = T-1 .. 0:
//   thread = (replica row, 32 hidden units): cell backward from gate activations, c_t, c_{t-1} and the
//   incoming dh / dc  ->  dz (4 x 32) written fp32 in place over the gates (operand of the weight-gradient
//   GEMMs) and as bf16 into the A tile [128 x 256];  tcgen05.mma  D[128 x 64] = dz . Wh^T  (K = 256, N = 64)
//   -> TMEM -> the thread's 32 columns = dh_{t-1} carry.  dc / dh carries stay in registers.
#define BW_KC 32            // 256 / 8 K-chunks
__global__ void pack_wht_kernel(const DDimsTC d, const float* __restrict__ P, __nv_bfloat16* __restrict__ Wt) {
  const int u = blockIdx.y;
  const float* Wh = P + d.off_wh + (int64_t)u * TC_H * TC_N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < BW_KC * TC_H; i += gridDim.x * blockDim.x) {
    const int kc = i / TC_H, n = i - kc * TC_H;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(Wh[(int64_t)n * TC_N + kc * 8 + e]);
    *reinterpret_cast<uint4*>(Wt + (((int64_t)u * BW_KC + kc) * TC_H + n) * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// B operand of the fused dX = dZ . Wx^T product: element (kc, n, e) = Wx[n][kc * 8 + e], n = fc-output column (< dx)
__global__ void pack_wxt_kernel(const DDimsTC d, const float* __restrict__ P, __nv_bfloat16* __restrict__ Wxt) {
  const int u = blockIdx.y;
  const float* Wx = P + d.off_wx + (int64_t)u * d.dx * TC_N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < BW_KC * d.dx; i += gridDim.x * blockDim.x) {
    const int kc = i / d.dx, n = i - kc * d.dx;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(Wx[(int64_t)n * TC_N + kc * 8 + e]);
    *reinterpret_cast<uint4*>(Wxt + (((int64_t)u * BW_KC + kc) * d.dx + n) * 8) = *reinterpret_cast<const uint4*>(v);
  }
}
struct BwdTC {
  const __nv_bfloat16* Wxt;  // optional [2A][32][dx][8] (tscl_pack_wxt): with dXb, dX = dZ . Wx^T is fused into the step
  __nv_bfloat16* dXb;        // optional [2A][T*Rc][dx] bf16
  const __nv_bfloat16* Wt;   // [2A][32][64][8]
  float* ZG;                 // [2A][T*Rc][256] gates in, dZ out
  const float* C;            // [2A][T*Rc][64]
  const float* dH;           // [2A][T*Rc][64]
  const float* c0;           // [2A][ld_state][64]
  const float* done;         // [T]
  int T;
  int64_t Rc, ld_state, r0;
  const __nv_bfloat16* Gb;   // optional: gate activations / c_t straight from the bf16 activation store
  const __nv_bfloat16* Cb;   //           ([2A][T*Rc][256] / [..][64]); when set, ZG is write-only and C is unused
  __nv_bfloat16* dZb;        // optional: dZ as bf16 [2A][T*Rc][256] (operand of the tensor-core weight-gradient kernels);
                             //           ZG may then be null (needs Gb / Cb)
  unsigned long long* prof;  // optional: 8 phase counters of the staged kernel (tscl_debug_bptt_prof)
};

// NT = 512: thread = (replica row, 16 hidden units), one CTA per SM.
// NT = 256: thread = (replica row, 32 hidden units) processed 8 at a time, TWO CTAs per SM (<= 128 registers): while one
//           tile waits for its per-step MMA / barrier the other one has its loads in flight.
template <int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1)
lstm_bwd_tc_kernel(const DDimsTC d, const BwdTC a) {
  constexpr int HPT = 8192 / NT;            // hidden units per thread
  constexpr int NSUB = HPT / 8;
  constexpr int NQ = NT / 128;              // hidden groups (threads per row)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool fuse_dx = a.dXb != nullptr;             // second product of the same dz tile: dX = dz . Wx^T (N = dx)
  const int dx = d.dx;
  unsigned char* sB = tc_smem;                       // 32 * 1024  : Wh^T image
  unsigned char* sA = sB + BW_KC * 1024;             // 32 * 2048  : dz tile
  uint64_t* sBar = reinterpret_cast<uint64_t*>(sA + BW_KC * 2048);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 1);
  unsigned char* sBx = sA + BW_KC * 2048 + 16;       // 32 * dx * 16 : Wx^T image (only when fuse_dx)
  const uint32_t bar = smem_u32(sBar);
  const uint32_t tmem_cols = fuse_dx ? 512u : 64u;   // D1 (dh) in columns 0..63, D2 (dX) in columns 64..64+dx
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *sTmem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_H >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  const uint32_t idesc_x = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(dx >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  const uint32_t aA = smem_u32(sA), aB = smem_u32(sB), aBx = smem_u32(sBx);
  const int64_t n_tiles = (a.Rc + TC_M - 1) / TC_M;
  const int64_t n_items = n_tiles * 2 * d.A;
  int cur_u = -1;
  uint32_t parity = 0;
  const int q = warp & 3, qt = warp >> 2;            // TMEM lane quadrant, hidden-unit group
  const int row = q * 32 + lane, jq = qt * HPT;
  auto bf8 = [](const uint4 v, float* o) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  };
  auto f8 = [](const float* p, float* o) {
    const float4 x = reinterpret_cast<const float4*>(p)[0], y = reinterpret_cast<const float4*>(p)[1];
    o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = y.x; o[5] = y.y; o[6] = y.z; o[7] = y.w;
  };
  for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int u = (int)(it / n_tiles);
    const int64_t r = (it - (int64_t)u * n_tiles) * TC_M + row;
    const bool valid = r < a.Rc;
    __syncthreads();
    if (u != cur_u) {
      cur_u = u;
      const uint4* src = reinterpret_cast<const uint4*>(a.Wt + (int64_t)u * BW_KC * TC_H * 8);
      uint4* dst = reinterpret_cast<uint4*>(sB);
      for (int i = tid; i < BW_KC * TC_H; i += NT) dst[i] = src[i];
      if (fuse_dx) {
        const uint4* sx = reinterpret_cast<const uint4*>(a.Wxt + (int64_t)u * BW_KC * dx * 8);
        uint4* dxs = reinterpret_cast<uint4*>(sBx);
        for (int i = tid; i < BW_KC * dx; i += NT) dxs[i] = sx[i];
      }
    }
    float dc[HPT], dhc[HPT];
#pragma unroll
    for (int e = 0; e < HPT; ++e) { dc[e] = 0.f; dhc[e] = 0.f; }
    for (int t = a.T - 1; t >= 0; --t) {
      const float keep = 1.0f - a.done[t];
      const int64_t m = ((int64_t)u * a.T + t) * a.Rc + (valid ? r : 0);
      if (t > 0 && valid) {      // pull step t-1's operands towards L2 while step t is being processed
        const int64_t mp = m - a.Rc;
        constexpr int LP = 4 / NQ;                 // 128-byte lines of a 512-byte row per thread
#pragma unroll
        for (int ln = 0; ln < LP; ++ln) {
          const int li = qt * LP + ln;             // 0..3
          if (a.Gb) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Gb + mp * TC_N + li * 64));
            if (li == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Cb + mp * TC_H));
            if (li == 1 && t > 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Cb + (mp - a.Rc) * TC_H));
          } else {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ZG + mp * TC_N + li * 64));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ZG + mp * TC_N + li * 64 + 32));
            if (li < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.C + mp * TC_H + li * 32));
          }
          if (li >= 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dH + mp * TC_H + (li - 2) * 32));
        }
      }
#pragma unroll
      for (int jb = 0; jb < NSUB; ++jb) {
        const int jo = jq + jb * 8;
        float gi[8], gf[8], go[8], gu[8], ct[8], cp[8], dh[8];
        if (valid) {
          if (a.Gb) {
            const __nv_bfloat16* zb = a.Gb + m * TC_N + jo;
            bf8(*reinterpret_cast<const uint4*>(zb), gi); bf8(*reinterpret_cast<const uint4*>(zb + 64), gf);
            bf8(*reinterpret_cast<const uint4*>(zb + 128), go); bf8(*reinterpret_cast<const uint4*>(zb + 192), gu);
            bf8(*reinterpret_cast<const uint4*>(a.Cb + m * TC_H + jo), ct);
            if (t > 0) bf8(*reinterpret_cast<const uint4*>(a.Cb + (m - a.Rc) * TC_H + jo), cp);
            if (t == 0) f8(a.c0 + ((int64_t)u * a.ld_state + a.r0 + r) * TC_H + jo, cp);
            f8(a.dH + m * TC_H + jo, dh);
          } else {
            const float* z = a.ZG + m * TC_N + jo;
            f8(z, gi); f8(z + 64, gf); f8(z + 128, go); f8(z + 192, gu);
            f8(a.C + m * TC_H + jo, ct);
            f8(t > 0 ? a.C + (m - a.Rc) * TC_H + jo : a.c0 + ((int64_t)u * a.ld_state + a.r0 + r) * TC_H + jo, cp);
            f8(a.dH + m * TC_H + jo, dh);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) cp[e] *= keep;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) { gi[e] = gf[e] = go[e] = gu[e] = ct[e] = cp[e] = dh[e] = 0.f; }
        }
        float dzi[8], dzf[8], dzo[8], dzu[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = jb * 8 + e;
          const float dht = dh[e] + dhc[k];
          const float tc = tanh_fast(ct[e]);
          const float dcc = dc[k] + dht * go[e] * (1.0f - tc * tc);
          dzi[e] = dcc * gu[e] * gi[e] * (1.0f - gi[e]);
          dzf[e] = dcc * cp[e] * gf[e] * (1.0f - gf[e]);
          dzo[e] = dht * tc * go[e] * (1.0f - go[e]);
          dzu[e] = dcc * gi[e] * (1.0f - gu[e] * gu[e]);
          dc[k] = dcc * gf[e] * keep;
        }
        const float* srcs[4] = {dzi, dzf, dzo, dzu};
        if (valid && a.ZG) {
          float4* z = reinterpret_cast<float4*>(a.ZG + m * TC_N + jo);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            z[g * 16] = make_float4(srcs[g][0], srcs[g][1], srcs[g][2], srcs[g][3]);
            z[g * 16 + 1] = make_float4(srcs[g][4], srcs[g][5], srcs[g][6], srcs[g][7]);
          }
        }
        // bf16 copies into the A tile: columns g*64 + jo .. +8  ->  chunk (g*64 + jo)/8, row `row`
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          __align__(16) __nv_bfloat16 v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(srcs[g][e]);
          *reinterpret_cast<uint4*>(sA + (size_t)((g * 64 + jo) >> 3) * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v);
          if (valid && a.dZb) *reinterpret_cast<uint4*>(a.dZb + m * TC_N + g * 64 + jo) = *reinterpret_cast<const uint4*>(v);
        }
      }
      const bool need_dh = keep != 0.f && t > 0;
      if (need_dh || fuse_dx) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
          if (lane == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (need_dh)
              for (int ks = 0; ks < 16; ++ks)
                umma_bf16(tmem, make_desc(aA + ks * 2 * 2048, 2048, 128), make_desc(aB + ks * 2 * 1024, 1024, 128), idesc,
                          ks > 0 ? 1u : 0u);
            if (fuse_dx)      // same A tile, B = Wx^T image: chunk stride dx * 16 B, 8-row groups 128 B apart
              for (int ks = 0; ks < 16; ++ks)
                umma_bf16(tmem + 64u, make_desc(aA + ks * 2 * 2048, 2048, 128),
                          make_desc(aBx + (uint32_t)(ks * 2 * dx * 16), (uint32_t)(dx * 16), 128), idesc_x, ks > 0 ? 1u : 0u);
            umma_commit(bar);
          }
          __syncwarp();
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (need_dh) {
          float dhp[HPT];
#pragma unroll
          for (int c16 = 0; c16 < HPT / 16; ++c16)
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(jq + c16 * 16), dhp + c16 * 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int e = 0; e < HPT; ++e) dhc[e] = dhp[e] * keep;
        } else {
#pragma unroll
          for (int e = 0; e < HPT; ++e) dhc[e] = 0.f;
        }
        if (fuse_dx) {        // this thread's share of the row: dx / NQ columns, 8 at a time -> bf16 -> one 16-byte store
          const int per = dx / NQ;
          for (int c8 = 0; c8 < per; c8 += 8) {
            float xv[8];
            tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 + qt * per + c8), xv);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (valid) {
              __align__(16) __nv_bfloat16 v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(xv[e]);
              *reinterpret_cast<uint4*>(a.dXb + m * dx + qt * per + c8) = *reinterpret_cast<const uint4*>(v);
            }
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < HPT; ++e) dhc[e] = 0.f;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
}

// ---------------------------------------------------------------------------------------------------
// BPTT, staged variant (the activation-store path of the training loop: gates / c from the bf16 store, dZ written as
// bf16).  Same arithmetic as lstm_bwd_tc_kernel<512>; what changes is how the operands reach the threads.  There a
// thread owns (row, 16 hidden units) and loads its 16-byte pieces straight from global memory, 128-512 B apart between
// the lanes of a warp (32 sectors per load instruction: lg_throttle 3.5 and long_scoreboard 10 warps per issue in
// profiles/r02).  Here each step's tile — gates [128 x 512 B], c_{t-1} [128 x 128 B], dH [128 x 256 B] — is fetched by
// coalesced 16-byte cp.async (one row segment per warp instruction) into XOR-swizzled shared memory one step AHEAD
// (issued as soon as every thread has taken step t's operands into registers, landing while step t's cell math, MMA
// and TMEM read-back run), and the threads pick their pieces from there without bank conflicts.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// TMA = true: the per-step operand tile (gates 64 KB, dH 32 KB, c 16 KB) arrives by SEVEN cp.async.bulk.tensor copies issued
// by one thread (hardware 128-byte swizzle = the pattern the readers use) instead of 14 cp.async per thread: the phase
// profile showed the issue of those 7168 16-byte copies blocking every warp for 7.9 k of the 17 k cycles of a step.
template <bool PROF, bool TMA>
__global__ void __launch_bounds__(512, 1)
lstm_bwd_tc_staged_kernel(const DDimsTC d, const BwdTC a, const __grid_constant__ CUtensorMap mapG,
                          const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapD) {
  constexpr int NT = 512, HPT = 16, NSUB = 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long bp[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc = 0;
#define BP_MARK(i) do { if (PROF && tid == 0) { const long long c_ = clock64(); bp[i] += c_ - pc; pc = c_; } } while (0)
  unsigned char* sB = tc_smem;                       // 32 KB : Wh^T image
  unsigned char* sA = sB + BW_KC * 1024;             // 64 KB : dz tile (A operand)
  unsigned char* sG = sA + BW_KC * 2048;             // 64 KB : gates [128][32 chunks ^ (row & 31)]   (1024-byte aligned)
                                                     //         TMA: 4 boxes [128 rows][8 chunks ^ (row & 7)], one per gate
  unsigned char* sC0 = sG + 128 * 512;               // 16 KB x 2 : c ring, [128][8 chunks ^ (row & 7)]
  unsigned char* sD = sC0 + 2 * 128 * 128;           // 32 KB : dH fp32 [128][16 chunks ^ (row & 15)]; TMA: 2 boxes of 32 floats
  uint64_t* sBar = reinterpret_cast<uint64_t*>(sD + 128 * 256);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 2);
  const uint32_t bar = smem_u32(sBar), ldbar = bar + 8;
  uint32_t ldpar = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar, 1); mbar_init(ldbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *sTmem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_H >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
  const uint32_t aG = smem_u32(sG), aC = smem_u32(sC0), aD = smem_u32(sD);
  const int64_t n_tiles = (a.Rc + TC_M - 1) / TC_M;
  const int64_t n_items = n_tiles * 2 * d.A;
  int cur_u = -1;
  uint32_t parity = 0;
  const int q = warp & 3, qt = warp >> 2;            // TMEM lane quadrant, hidden-unit group (16 units)
  const int row = q * 32 + lane, jq = qt * HPT;
  auto bf8 = [](const uint4 v, float* o) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  };
  for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int u = (int)(it / n_tiles);
    const int64_t rt0 = (it - (int64_t)u * n_tiles) * TC_M;       // first replica row of the tile
    const int64_t r = rt0 + row;
    const bool valid = r < a.Rc;
    __syncthreads();
    if (u != cur_u) {
      cur_u = u;
      const uint4* src = reinterpret_cast<const uint4*>(a.Wt + (int64_t)u * BW_KC * TC_H * 8);
      uint4* dst = reinterpret_cast<uint4*>(sB);
      for (int i = tid; i < BW_KC * TC_H; i += NT) dst[i] = src[i];
    }
    // coalesced fetch of one step's tile: rows beyond Rc are clamped to a valid row (their results are never stored)
    auto fetch = [&](int t, bool with_c_t) {
      const int64_t mb = ((int64_t)u * a.T + t) * a.Rc;            // row index of replica 0 at step t
      if constexpr (TMA) {
        if (tid == 0) {       // rows past the end of the tensors are zero-filled; rows past Rc belong to other tiles and are never stored
          auto tma2d = [&](uint32_t dst, const CUtensorMap* mp, int x, int64_t y) {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(mp)), "r"(x), "r"((int)y), "r"(ldbar) : "memory");
          };
          const int64_t y = mb + rt0;
          mbar_expect_tx(ldbar, (uint32_t)(65536 + 32768 + (with_c_t ? 16384 : 0) + (t > 0 ? 16384 : 0)));
#pragma unroll
          for (int g = 0; g < 4; ++g) tma2d(aG + g * 16384, &mapG, g * 64, y);
          tma2d(aD, &mapD, 0, y); tma2d(aD + 16384, &mapD, 32, y);
          if (with_c_t) tma2d(aC + (t & 1) * 16384, &mapC, 0, y);
          if (t > 0) tma2d(aC + ((t - 1) & 1) * 16384, &mapC, 0, y - a.Rc);
        }
        return;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {                                // gates: 128 rows x 32 chunks
        const int id = i * NT + tid, rw = id >> 5, c = id & 31;
        const int64_t rr = rt0 + rw < a.Rc ? rt0 + rw : a.Rc - 1;
        cp_async16(aG + rw * 512 + ((c ^ (rw & 31)) << 4), a.Gb + (mb + rr) * TC_N + c * 8);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {                                // dH fp32: 128 rows x 16 chunks
        const int id = i * NT + tid, rw = id >> 4, c = id & 15;
        const int64_t rr = rt0 + rw < a.Rc ? rt0 + rw : a.Rc - 1;
        cp_async16(aD + rw * 256 + ((c ^ (rw & 15)) << 4), a.dH + (mb + rr) * TC_H + c * 4);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {                                // c_{t-1} (and c_t for the first step): 128 rows x 8 chunks
        const int id = i * NT + tid, rw = id >> 3, c = id & 7;
        const int64_t rr = rt0 + rw < a.Rc ? rt0 + rw : a.Rc - 1;
        const uint32_t off = rw * 128 + ((c ^ (rw & 7)) << 4);
        if (with_c_t) cp_async16(aC + (t & 1) * 16384 + off, a.Cb + (mb + rr) * TC_H + c * 8);
        if (t > 0) cp_async16(aC + ((t - 1) & 1) * 16384 + off, a.Cb + (mb - a.Rc + rr) * TC_H + c * 8);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(a.T - 1, true);
    float dc[HPT], dhc[HPT];
#pragma unroll
    for (int e = 0; e < HPT; ++e) { dc[e] = 0.f; dhc[e] = 0.f; }
    if (PROF && tid == 0) pc = clock64();
    for (int t = a.T - 1; t >= 0; --t) {
      const float keep = 1.0f - a.done[t];
      const int64_t m = ((int64_t)u * a.T + t) * a.Rc + (valid ? r : 0);
      if (!TMA && t > 1) {      // pull step t-2's operands towards L2: they are fetched into shared memory during step t-1
        const int64_t mp = ((int64_t)u * a.T + t - 2) * a.Rc + (valid ? r : 0);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Gb + mp * TC_N + qt * 64));
        if (qt == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Cb + mp * TC_H));
        if (qt >= 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dH + mp * TC_H + (qt - 2) * 32));
      }
      if constexpr (TMA) {
        mbar_wait(ldbar, ldpar); ldpar ^= 1;          // step t's tile has landed (complete_tx of all its copies)
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                            // step t's tile is in shared memory
      }
      BP_MARK(0);                                   // waiting for step t's operands
      const unsigned char* cT = sC0 + (t & 1) * 16384;
      const unsigned char* cP = sC0 + ((t - 1) & 1) * 16384;
      uint4 zpk[4][NSUB];
#pragma unroll
      for (int jb = 0; jb < NSUB; ++jb) {
        const int jo = jq + jb * 8;
        float gi[8], gf[8], go[8], gu[8], ct[8], cp[8], dh[8];
        {
          const int cg = jo >> 3, sw = row & 31;     // chunk of 8 hidden units inside each 64-wide gate block
          if constexpr (TMA) {
            const unsigned char* gp = sG + row * 128 + ((cg ^ (row & 7)) << 4);
            bf8(*reinterpret_cast<const uint4*>(gp), gi);
            bf8(*reinterpret_cast<const uint4*>(gp + 16384), gf);
            bf8(*reinterpret_cast<const uint4*>(gp + 32768), go);
            bf8(*reinterpret_cast<const uint4*>(gp + 49152), gu);
          } else {
          bf8(*reinterpret_cast<const uint4*>(sG + row * 512 + (((cg) ^ sw) << 4)), gi);
          bf8(*reinterpret_cast<const uint4*>(sG + row * 512 + (((8 + cg) ^ sw) << 4)), gf);
          bf8(*reinterpret_cast<const uint4*>(sG + row * 512 + (((16 + cg) ^ sw) << 4)), go);
          bf8(*reinterpret_cast<const uint4*>(sG + row * 512 + (((24 + cg) ^ sw) << 4)), gu);
          }
          bf8(*reinterpret_cast<const uint4*>(cT + row * 128 + ((cg ^ (row & 7)) << 4)), ct);
          if (t > 0) bf8(*reinterpret_cast<const uint4*>(cP + row * 128 + ((cg ^ (row & 7)) << 4)), cp);
          else if (valid) {
            const float4 x = reinterpret_cast<const float4*>(a.c0 + ((int64_t)u * a.ld_state + a.r0 + r) * TC_H + jo)[0];
            const float4 y = reinterpret_cast<const float4*>(a.c0 + ((int64_t)u * a.ld_state + a.r0 + r) * TC_H + jo)[1];
            cp[0] = x.x; cp[1] = x.y; cp[2] = x.z; cp[3] = x.w; cp[4] = y.x; cp[5] = y.y; cp[6] = y.z; cp[7] = y.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) cp[e] = 0.f;
          }
          const int cd = jo >> 2;                    // dH: 4 floats per chunk
          const float4 d0 = TMA ? *reinterpret_cast<const float4*>(sD + (cd >> 3) * 16384 + row * 128 + (((cd & 7) ^ (row & 7)) << 4))
                                : *reinterpret_cast<const float4*>(sD + row * 256 + ((cd ^ (row & 15)) << 4));
          const float4 d1 = TMA ? *reinterpret_cast<const float4*>(sD + (cd >> 3) * 16384 + row * 128 + ((((cd + 1) & 7) ^ (row & 7)) << 4))
                                : *reinterpret_cast<const float4*>(sD + row * 256 + (((cd + 1) ^ (row & 15)) << 4));
          dh[0] = d0.x; dh[1] = d0.y; dh[2] = d0.z; dh[3] = d0.w; dh[4] = d1.x; dh[5] = d1.y; dh[6] = d1.z; dh[7] = d1.w;
#pragma unroll
          for (int e = 0; e < 8; ++e) cp[e] *= keep;
          if (!valid) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { gi[e] = gf[e] = go[e] = gu[e] = ct[e] = cp[e] = dh[e] = 0.f; }
          }
        }
        if (jb == NSUB - 1) {      // every thread holds its last operands: the staging buffers can take step t-1
          BP_MARK(5);              // shared memory -> registers (both sub-batches) + first sub-batch's math and stores
          __syncthreads();
          BP_MARK(6);              // barrier: staging buffers free
          if (t > 0) fetch(t - 1, false);
          BP_MARK(7);              // cp.async issue of step t-1
        }
        float dzi[8], dzf[8], dzo[8], dzu[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = jb * 8 + e;
          const float dht = dh[e] + dhc[k];
          const float tc = tanh_fast(ct[e]);
          const float dcc = dc[k] + dht * go[e] * (1.0f - tc * tc);
          dzi[e] = dcc * gu[e] * gi[e] * (1.0f - gi[e]);
          dzf[e] = dcc * cp[e] * gf[e] * (1.0f - gf[e]);
          dzo[e] = dht * tc * go[e] * (1.0f - go[e]);
          dzu[e] = dcc * gi[e] * (1.0f - gu[e] * gu[e]);
          dc[k] = dcc * gf[e] * keep;
        }
        const float* srcs[4] = {dzi, dzf, dzo, dzu};
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          __align__(16) __nv_bfloat16 v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(srcs[g][e]);
          *reinterpret_cast<uint4*>(sA + (size_t)((g * 64 + jo) >> 3) * 2048 + row * 16) = *reinterpret_cast<const uint4*>(v);
          zpk[g][jb] = *reinterpret_cast<const uint4*>(v);
        }
      }
      if (valid) {      // dZ of this thread's 16 hidden units: one 256-bit store (a full sector) per gate
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 x = zpk[g][0], y = zpk[g][1];
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(a.dZb + m * TC_N + g * 64 + jq), "r"(x.x),
                       "r"(x.y), "r"(x.z), "r"(x.w), "r"(y.x), "r"(y.y), "r"(y.z), "r"(y.w) : "memory");
        }
      }
      BP_MARK(1);                                   // second sub-batch: cell backward, dZ stores
      if (keep != 0.f && t > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        BP_MARK(2);                                 // barrier before the MMA
        if (warp == 0) {
          if (lane == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int ks = 0; ks < 16; ++ks)
              umma_bf16(tmem, make_desc(aA + ks * 2 * 2048, 2048, 128), make_desc(aB + ks * 2 * 1024, 1024, 128), idesc,
                        ks > 0 ? 1u : 0u);
            umma_commit(bar);
          }
          __syncwarp();
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        BP_MARK(3);                                 // MMA issue + commit + wait
        float dhp[HPT];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)jq, dhp);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < HPT; ++e) dhc[e] = dhp[e] * keep;
        BP_MARK(4);                                 // TMEM read-back
      } else {
#pragma unroll
        for (int e = 0; e < HPT; ++e) dhc[e] = 0.f;
      }
    }
  }
  if (PROF && tid == 0)
    for (int i = 0; i < 8; ++i) atomicAdd(a.prof + i, (unsigned long long)bp[i]);
#undef BP_MARK
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

// 2-D tiled tensor map (row-major [rows][cols], box = box_cols x box_rows, inner box = 128 bytes, 128-byte swizzle) through the
// driver entry point (no link-time dependency on libcuda)
static bool make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t rows, uint64_t cols,
                         uint32_t box_cols, uint32_t box_rows) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = []() -> encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (encode_fn)p;
  }();
  if (!fn || !base) return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * (uint64_t)elem_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t es[2] = {1, 1};
  return fn(m, dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int tscl_pack_wxt(tscl_handle* h, const float* params, void* wxt_bf16, void* stream) {
  if (!h || !params || !wxt_bf16) return tsc_set_error("tscl_pack_wxt: bad argument");
  PCK(cudaSetDevice(tscl_device_of(h)));
  const DDimsTC& d = *tscl_dims_of(h);
  pack_wxt_kernel<<<dim3(8, 2 * d.A), 256, 0, (cudaStream_t)stream>>>(d, params, (__nv_bfloat16*)wxt_bf16);
  PCK(cudaGetLastError());
  return 0;
}

extern "C" int tscl_pack_wht(tscl_handle* h, const float* params, void* wt_bf16, void* stream) {
  if (!h || !params || !wt_bf16) return tsc_set_error("tscl_pack_wht: bad argument");
  PCK(cudaSetDevice(tscl_device_of(h)));
  const DDimsTC& d = *tscl_dims_of(h);
  pack_wht_kernel<<<dim3(2, 2 * d.A), 256, 0, (cudaStream_t)stream>>>(d, params, (__nv_bfloat16*)wt_bf16);
  PCK(cudaGetLastError());
  return 0;
}

extern "C" int tscl_lstm_seq_bwd_tc(tscl_handle* h, const void* wt_bf16, float* ZG, const float* C, const float* dH,
                                    const float* c0, const float* done, int32_t T, int64_t Rc, int64_t ld_state,
                                    int64_t r0, const void* gates_bf16, const void* c_bf16, void* dz_bf16, void* stream) {
  return tscl_lstm_seq_bwd_tc_dx(h, wt_bf16, ZG, C, dH, c0, done, T, Rc, ld_state, r0, gates_bf16, c_bf16, dz_bf16, nullptr,
                                 nullptr, stream);
}

extern "C" int tscl_lstm_seq_bwd_tc_dx(tscl_handle* h, const void* wt_bf16, float* ZG, const float* C, const float* dH,
                                       const float* c0, const float* done, int32_t T, int64_t Rc, int64_t ld_state,
                                       int64_t r0, const void* gates_bf16, const void* c_bf16, void* dz_bf16,
                                       const void* wxt_bf16, void* dx_bf16, void* stream) {
  if (!h || !wt_bf16 || T <= 0 || Rc <= 0) return tsc_set_error("tscl_lstm_seq_bwd_tc: bad argument");
  if ((wxt_bf16 == nullptr) != (dx_bf16 == nullptr)) return tsc_set_error("tscl_lstm_seq_bwd_tc_dx: wxt_bf16 and dx_bf16 go together");
  if (!ZG && !(gates_bf16 && c_bf16 && dz_bf16)) return tsc_set_error("tscl_lstm_seq_bwd_tc: ZG may be null only with gates_bf16, c_bf16 and dz_bf16");
  if ((gates_bf16 == nullptr) != (c_bf16 == nullptr)) return tsc_set_error("tscl_lstm_seq_bwd_tc: gates_bf16 and c_bf16 go together");
  PCK(cudaSetDevice(tscl_device_of(h)));
  const DDimsTC& d = *tscl_dims_of(h);
  if (dx_bf16 && (d.dx % 32 != 0 || d.dx > 256)) return tsc_set_error("tscl_lstm_seq_bwd_tc_dx: dx must be a multiple of 32, <= 256");
  const size_t smem_max = BW_KC * 1024 + BW_KC * 2048 + 16 + (size_t)BW_KC * 256 * 16;
  const size_t smem = BW_KC * 1024 + BW_KC * 2048 + 16 + (dx_bf16 ? (size_t)BW_KC * d.dx * 16 : 0);
  static int attr_dev = -1;
  if (attr_dev != tscl_device_of(h)) {
    PCK(cudaFuncSetAttribute(lstm_bwd_tc_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    PCK(cudaFuncSetAttribute(lstm_bwd_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    attr_dev = tscl_device_of(h);
  }
  int n_sm = 0;
  PCK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, tscl_device_of(h)));
  const int64_t n_items = ((Rc + TC_M - 1) / TC_M) * 2 * d.A;
  // 96 KB smem: two 512-thread CTAs per SM when registers allow; the fused-dX variant (up to 211 KB, 512 TMEM columns) is
  // one CTA per SM
  const int64_t max_ctas = dx_bf16 ? n_sm : 2 * n_sm;
  const int grid = (int)(n_items < max_ctas ? n_items : max_ctas);
  BwdTC a;
  a.Wt = (const __nv_bfloat16*)wt_bf16; a.ZG = ZG; a.C = C; a.dH = dH; a.c0 = c0; a.done = done; a.T = T; a.Rc = Rc;
  a.ld_state = ld_state; a.r0 = r0; a.Gb = (const __nv_bfloat16*)gates_bf16; a.Cb = (const __nv_bfloat16*)c_bf16; a.dZb = (__nv_bfloat16*)dz_bf16;
  a.Wxt = (const __nv_bfloat16*)wxt_bf16; a.dXb = (__nv_bfloat16*)dx_bf16;
  // measured (R = 8192, 1 x B200): the one-CTA-per-SM 512-thread variant 2.097 ms per control step, the two-CTA 256-thread
  // variant 2.144 ms; TSC_BPTT_THREADS=256 selects the latter for experiments
  static const int bw_threads = []() { const char* e = getenv("TSC_BPTT_THREADS"); return e && atoi(e) == 256 ? 256 : 512; }();
  // staged variant (coalesced cp.async into swizzled shared memory one step ahead): store path without fused dX
  static const int bw_staged = []() { const char* e = getenv("TSC_BPTT_STAGED"); return e ? atoi(e) : 1; }();
  if (bw_staged && bw_threads == 512 && a.Gb && a.Cb && a.dZb && !a.ZG && !a.dXb) {
    const size_t smem_s = BW_KC * 1024 + BW_KC * 2048 + 128 * 512 + 2 * 128 * 128 + 128 * 256 + 32;
    static int attr_s = -1;
    if (attr_s != tscl_device_of(h)) {
      PCK(cudaFuncSetAttribute(lstm_bwd_tc_staged_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
      PCK(cudaFuncSetAttribute(lstm_bwd_tc_staged_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
      PCK(cudaFuncSetAttribute(lstm_bwd_tc_staged_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
      PCK(cudaFuncSetAttribute(lstm_bwd_tc_staged_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
      attr_s = tscl_device_of(h);
    }
    const int grid_s = (int)(n_items < n_sm ? n_items : n_sm);
    a.prof = g_bptt_prof;
    // operand tiles by TMA (tensor maps over this chunk's gate / c / dH arrays, 128-byte swizzle); TSC_BPTT_TMA=0: cp.async
    static const int bw_tma = []() { const char* e = getenv("TSC_BPTT_TMA"); return e ? atoi(e) : 1; }();
    CUtensorMap mG, mC, mD;
    memset(&mG, 0, sizeof(mG)); memset(&mC, 0, sizeof(mC)); memset(&mD, 0, sizeof(mD));
    const uint64_t rows = (uint64_t)2 * d.A * T * Rc;
    const bool tma = bw_tma && rows < (1ull << 31) &&
                     make_tmap_2d(&mG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gates_bf16, rows, TC_N, 64, 128) &&
                     make_tmap_2d(&mC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c_bf16, rows, TC_H, 64, 128) &&
                     make_tmap_2d(&mD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dH, rows, TC_H, 32, 128);
    if (tma) {
      if (a.prof) lstm_bwd_tc_staged_kernel<true, true><<<grid_s, 512, smem_s, (cudaStream_t)stream>>>(d, a, mG, mC, mD);
      else lstm_bwd_tc_staged_kernel<false, true><<<grid_s, 512, smem_s, (cudaStream_t)stream>>>(d, a, mG, mC, mD);
    } else {
      if (a.prof) lstm_bwd_tc_staged_kernel<true, false><<<grid_s, 512, smem_s, (cudaStream_t)stream>>>(d, a, mG, mC, mD);
      else lstm_bwd_tc_staged_kernel<false, false><<<grid_s, 512, smem_s, (cudaStream_t)stream>>>(d, a, mG, mC, mD);
    }
  } else if (bw_threads == 512) lstm_bwd_tc_kernel<512><<<grid, 512, smem, (cudaStream_t)stream>>>(d, a);
  else lstm_bwd_tc_kernel<256><<<grid, 256, smem, (cudaStream_t)stream>>>(d, a);
  PCK(cudaGetLastError());
  return 0;
}

// ===================================================================================================
// fc front-end weight gradients on the tensor cores (replaces fc_bwd_kernel of tsc_learn.cu).
//   dW[k][c] = sum_m In[m][k] * dXm[m][c],   dXm = dX * (X > 0),   m = (t, replica) rows of one chunk
// is a GEMM whose reduction index is the ROW index, so both operands are staged MN-major: the row-major global
// data lands as [col/8][128 rows][8 cols] bf16 without a transposition, and
//   D[dX column (two M = 128 halves)][64 input slots] += A^T B      (tcgen05.mma, a_major = b_major = MN, K = 128 rows)
// accumulates in TMEM over all tiles of a unit.  A spare input slot holds 1.0: its D column is the bias gradient.
// Persistent: CTA b owns the tile range [b NT / grid, (b + 1) NT / grid) of the (unit, tile) list and flushes its
// accumulator with atomics whenever the unit changes (<= 3 flushes per CTA).
#define FBT_ROWS 128
#define FBT_SBO 2064                       // chunk stride: 128 rows * 16 B + 16 B pad (conflict-free transposing stores)
#define FBT_A_BYTES (32 * FBT_SBO)
#define FBT_B_BYTES (8 * FBT_SBO)
#define FBT_STAGE (FBT_A_BYTES + FBT_B_BYTES)
#define FBT_THREADS 512
struct FcBwdTC {
  const float* obs;            // rows as tscl_fc_embed
  const float* X;              // [2A][M][dx] fp32 activations, or
  const __nv_bfloat16* Xb;     // [2A][M][dx] bf16 activations (one chunk of the activation store)
  const float* dX;             // [2A][M][dx] fp32, or
  const __nv_bfloat16* dXb;    // [2A][M][dx] bf16
  float* G;
  int64_t M, rows_per_t, stride_t;
  int variant;                 // 1: LBO/SBO swapped (descriptor diagnosis)
};

__global__ void __launch_bounds__(FBT_THREADS, 1)
fc_bwd_tc_kernel(const DDimsTC d, const FcBwdTC a) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(tc_smem + 2 * FBT_STAGE);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 2);
  const uint32_t bar0 = smem_u32(sBar);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *sTmem;
  // bf16 x bf16 -> f32, A and B MN-major, N = 64, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  const uint32_t lbo = a.variant ? FBT_SBO : 128, sbo = a.variant ? 128 : FBT_SBO;
  const int dx = d.dx, ng = dx >> 3, n_items = FBT_ROWS * ng;
  const uint32_t ng_magic = (1u << 20) / (uint32_t)ng + 1u;      // i / ng == (i * magic) >> 20 for i < 4096, ng <= 32
  const int64_t tpu = (a.M + FBT_ROWS - 1) / FBT_ROWS;
  const int64_t NT = tpu * 2 * d.A;
  const int64_t j0 = NT * blockIdx.x / gridDim.x, j1 = NT * (blockIdx.x + 1) / gridDim.x;
  uint32_t ph0 = 0, ph1 = 0;
  bool pend0 = false, pend1 = false, first = true;
  int cur_u = -1, nw = 0, nt = 0, nf = 0, ooff = 0;
  int src[8];                                  // observation index of each of this thread's 8 input slots (-1 none, -2 one)
  const int bc = tid & 7;                      // this thread's B chunk (8 input slots)

  auto flush = [&](int u) {
    if (pend0) { mbar_wait(bar0, ph0); ph0 ^= 1; pend0 = false; }
    if (pend1) { mbar_wait(bar0 + 8, ph1); ph1 ^= 1; pend1 = false; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, cgp = warp >> 2, mh = cgp >> 1, k0 = (cgp & 1) * 32;
    const int c = mh * 128 + q * 32 + lane;
    float v[32];
    const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * 64 + k0);
    tmem_ld16(tb, v); tmem_ld16(tb + 16, v + 16);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (c < dx) {
      int64_t wo, bo; int ld, cc, s0, n;
      if (c < d.fw) { wo = d.off_fcw_w[u]; bo = d.off_fcw_b[u]; ld = d.fw; cc = c; s0 = 0; n = nw; }
      else if (c < d.fw + d.ff) { wo = d.off_fcf_w[u]; bo = d.off_fcf_b[u]; ld = d.ff; cc = c - d.fw; s0 = d.kw; n = nf; }
      else { wo = d.off_fct_w[u]; bo = d.off_fct_b[u]; ld = d.ft; cc = c - d.fw - d.ff; s0 = d.kw + TC_KF; n = nt; }
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int slot = k0 + e, kin = slot - s0;
        if (kin >= 0 && kin < n) atomicAdd(&a.G[wo + (int64_t)kin * ld + cc], v[e]);
        if (slot == d.ones_slot) atomicAdd(&a.G[bo + cc], v[e]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  };

  for (int64_t j = j0; j < j1; ++j) {
    const int u = (int)(j / tpu);
    const int64_t m0 = (j - (int64_t)u * tpu) * FBT_ROWS;
    const int s = (int)((j - j0) & 1);
    if (u != cur_u) {
      if (cur_u >= 0) flush(cur_u);
      cur_u = u; first = true;
      const int ag = u >> 1;
      nw = d.n_wave[ag]; nt = d.n_wait[ag]; nf = d.ff > 0 ? d.n_fp[ag] : 0; ooff = d.obs_off[ag];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int slot = bc * 8 + e;
        int sidx = -1;
        if (slot < d.kw) { if (slot < nw) sidx = slot; }
        else if (slot < d.kw + TC_KF) { if (slot - d.kw < nf) sidx = nw + nt + (slot - d.kw); }
        else { if (slot - d.kw - TC_KF < nt) sidx = nw + (slot - d.kw - TC_KF); }
        if (slot == d.ones_slot) sidx = -2;
        src[e] = sidx;
      }
    }
    // the MMAs that read stage s two tiles ago must have drained it
    if (s == 0) { if (pend0) { mbar_wait(bar0, ph0); ph0 ^= 1; pend0 = false; } }
    else { if (pend1) { mbar_wait(bar0 + 8, ph1); ph1 ^= 1; pend1 = false; } }
    unsigned char* sA = tc_smem + (size_t)s * FBT_STAGE;
    unsigned char* sB = sA + FBT_A_BYTES;
    const int rows_valid = (a.M - m0) < FBT_ROWS ? (int)(a.M - m0) : FBT_ROWS;
    const int items_valid = rows_valid * ng;
    const int64_t base = ((int64_t)u * a.M + m0) * dx;
    if (j + 1 < j1) {      // pull the next tile towards L2 while this one is converted and multiplied
      const int un = (int)((j + 1) / tpu);
      const int64_t m0n = (j + 1 - (int64_t)un * tpu) * FBT_ROWS;
      const int rvn = (a.M - m0n) < FBT_ROWS ? (int)(a.M - m0n) : FBT_ROWS;
      const int64_t basen = ((int64_t)un * a.M + m0n) * dx;
      const int64_t nbytes = (int64_t)rvn * dx * 4;
      if (a.dXb) {
        const char* pd = reinterpret_cast<const char*>(a.dXb + basen);
        for (int64_t o = (int64_t)tid * 128; o < nbytes / 2; o += FBT_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pd + o));
      } else {
        const char* pd = reinterpret_cast<const char*>(a.dX + basen);
        for (int64_t o = (int64_t)tid * 128; o < nbytes; o += FBT_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pd + o));
      }
      if (a.Xb) {
        const char* px = reinterpret_cast<const char*>(a.Xb + basen);
        for (int64_t o = (int64_t)tid * 128; o < nbytes / 2; o += FBT_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + o));
      }
      if (tid < rvn) {
        const int64_t m = m0n + tid;
        const float* op = a.obs + (m / a.rows_per_t) * a.stride_t + (m % a.rows_per_t) * d.n_obs + d.obs_off[un >> 1];
        asm volatile("prefetch.global.L2 [%0];" ::"l"(op));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(op + 32));
      }
    }
    // ---- B loads first (observation slice of this thread's 8 input slots, 2 rows): their latency overlaps the A staging ----
    float ov[(FBT_ROWS * 8) / FBT_THREADS][8];
    {
      const int64_t tq = m0 / a.rows_per_t, rem0 = m0 - tq * a.rows_per_t;      // one division per tile
#pragma unroll
      for (int r = 0; r < (FBT_ROWS * 8) / FBT_THREADS; ++r) {
        const int row = (r * FBT_THREADS + tid) >> 3;
        if (row < rows_valid) {
          int64_t tt = tq, rem = rem0 + row;
          while (rem >= a.rows_per_t) { rem -= a.rows_per_t; ++tt; }
          const float* op = a.obs + tt * a.stride_t + rem * d.n_obs + ooff;
#pragma unroll
          for (int e = 0; e < 8; ++e) ov[r][e] = src[e] >= 0 ? __ldg(op + src[e]) : (src[e] == -2 ? 1.0f : 0.f);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) ov[r][e] = 0.f;
        }
      }
    }
    if (a.dXb && a.Xb) {
      // ---- A, bf16 in / bf16 out: all loads of the tile in flight at once, relu mask as packed 16-bit integer ops ----
      uint4 gb[8], xm[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = k * FBT_THREADS + tid;
        gb[k] = make_uint4(0, 0, 0, 0); xm[k] = make_uint4(0, 0, 0, 0);
        if (i < items_valid) {
          gb[k] = __ldg(reinterpret_cast<const uint4*>(a.dXb + base) + i);
          xm[k] = __ldg(reinterpret_cast<const uint4*>(a.Xb + base) + i);
        }
      }
      auto keep2 = [](uint32_t x) -> uint32_t {      // 0xFFFF per 16-bit half where the bf16 value is > 0
        const uint32_t nz = ((x & 0x7FFF7FFFu) + 0x7FFF7FFFu) & ~x & 0x80008000u;
        return (nz >> 15) * 0xFFFFu;
      };
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = k * FBT_THREADS + tid;
        if (i < n_items) {
          const uint4 o = make_uint4(gb[k].x & keep2(xm[k].x), gb[k].y & keep2(xm[k].y), gb[k].z & keep2(xm[k].z),
                                     gb[k].w & keep2(xm[k].w));
          const int row = (int)(((uint32_t)i * ng_magic) >> 20), cg = i - row * ng;
          *reinterpret_cast<uint4*>(sA + (size_t)cg * FBT_SBO + row * 16) = o;
        }
      }
    } else {
      // ---- A: masked dX, 8 columns (one 16 B chunk row) per item; items are contiguous in global memory ----
      for (int ib = 0; ib < n_items; ib += 4 * FBT_THREADS) {
        float4 g0[4], g1[4];
        uint4 xb[4];
        float4 x0[4], x1[4];
  #pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = ib + r * FBT_THREADS + tid;
          g0[r] = g1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
          xb[r] = make_uint4(0, 0, 0, 0);
          x0[r] = x1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < items_valid) {
            if (a.dXb) {
              const uint4 gb = __ldg(reinterpret_cast<const uint4*>(a.dXb + base) + i);
              g0[r] = make_float4(__uint_as_float(gb.x << 16), __uint_as_float(gb.x & 0xffff0000u), __uint_as_float(gb.y << 16),
                                  __uint_as_float(gb.y & 0xffff0000u));
              g1[r] = make_float4(__uint_as_float(gb.z << 16), __uint_as_float(gb.z & 0xffff0000u), __uint_as_float(gb.w << 16),
                                  __uint_as_float(gb.w & 0xffff0000u));
            } else {
              const float4* gp = reinterpret_cast<const float4*>(a.dX + base) + 2 * (int64_t)i;
              g0[r] = __ldg(gp); g1[r] = __ldg(gp + 1);
            }
            if (a.Xb) xb[r] = __ldg(reinterpret_cast<const uint4*>(a.Xb + base) + i);
            else {
              const float4* xp = reinterpret_cast<const float4*>(a.X + base) + 2 * (int64_t)i;
              x0[r] = __ldg(xp); x1[r] = __ldg(xp + 1);
            }
          }
        }
  #pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = ib + r * FBT_THREADS + tid;
          if (i < n_items) {
            const float gv[8] = {g0[r].x, g0[r].y, g0[r].z, g0[r].w, g1[r].x, g1[r].y, g1[r].z, g1[r].w};
            bool pos[8];
            if (a.Xb) {
              const uint32_t w[4] = {xb[r].x, xb[r].y, xb[r].z, xb[r].w};
  #pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t lo = w[e] & 0xffffu, hi = w[e] >> 16;
                pos[2 * e] = (lo & 0x8000u) == 0 && (lo & 0x7fffu) != 0;
                pos[2 * e + 1] = (hi & 0x8000u) == 0 && (hi & 0x7fffu) != 0;
              }
            } else {
              const float xv[8] = {x0[r].x, x0[r].y, x0[r].z, x0[r].w, x1[r].x, x1[r].y, x1[r].z, x1[r].w};
  #pragma unroll
              for (int e = 0; e < 8; ++e) pos[e] = xv[e] > 0.f;
            }
            __align__(16) __nv_bfloat16 o[8];
  #pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float2bfloat16_rn(pos[e] ? gv[e] : 0.f);
            const int row = i / ng, cg = i - row * ng;
            *reinterpret_cast<uint4*>(sA + (size_t)cg * FBT_SBO + row * 16) = *reinterpret_cast<const uint4*>(o);
          }
        }
      }
    }
    // ---- B: the unit's observation slice scattered into the 64 input slots ----
#pragma unroll
    for (int r = 0; r < (FBT_ROWS * 8) / FBT_THREADS; ++r) {
      const int row = (r * FBT_THREADS + tid) >> 3;
      __align__(16) __nv_bfloat16 o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = __float2bfloat16_rn(ov[r][e]);
      *reinterpret_cast<uint4*>(sB + (size_t)bc * FBT_SBO + row * 16) = *reinterpret_cast<const uint4*>(o);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
#pragma unroll
      for (int mh = 0; mh < 2; ++mh)
        for (int ks = 0; ks < FBT_ROWS / 16; ++ks)
          umma_bf16(tmem + mh * 64, make_desc(aA + mh * 16 * FBT_SBO + ks * 256, lbo, sbo),
                    make_desc(aB + ks * 256, lbo, sbo), idesc, (first && ks == 0) ? 0u : 1u);
      umma_commit(bar0 + 8 * s);
    }
    if (s == 0) pend0 = true; else pend1 = true;
    first = false;
  }
  if (cur_u >= 0) flush(cur_u);
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

extern "C" int tscl_fc_bwd_tc(tscl_handle* h, const float* obs, const float* X, const void* x_bf16, const float* dX,
                              const void* dx_bf16, int64_t M, int64_t rows_per_t, int64_t stride_t, float* grads,
                              int32_t variant, void* stream) {
  if (!h || !obs || (!X && !x_bf16) || (!dX && !dx_bf16) || !grads || M <= 0) return tsc_set_error("tscl_fc_bwd_tc: bad argument");
  PCK(cudaSetDevice(tscl_device_of(h)));
  const DDimsTC& d = *tscl_dims_of(h);
  if ((d.dx % 8) != 0 || d.dx > 256) return tsc_set_error("tscl_fc_bwd_tc: dx must be a multiple of 8, <= 256");
  if (d.kw == 0 || d.ones_slot < 0) return tsc_set_error("tscl_fc_bwd_tc: no free input slot for the bias column (use tscl_fc_bwd)");
  const size_t smem = 2 * FBT_STAGE + 32;
  static int attr_dev = -1;
  if (attr_dev != tscl_device_of(h)) {
    PCK(cudaFuncSetAttribute(fc_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev = tscl_device_of(h);
  }
  int n_sm = 0;
  PCK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, tscl_device_of(h)));
  const int64_t NT = ((M + FBT_ROWS - 1) / FBT_ROWS) * 2 * d.A;
  const int grid = (int)(NT < n_sm ? NT : n_sm);
  FcBwdTC a;
  a.obs = obs; a.X = X; a.Xb = (const __nv_bfloat16*)x_bf16; a.dX = dX; a.dXb = (const __nv_bfloat16*)dx_bf16; a.G = grads; a.M = M;
  a.rows_per_t = rows_per_t;
  a.stride_t = stride_t; a.variant = variant;
  fc_bwd_tc_kernel<<<grid, FBT_THREADS, smem, (cudaStream_t)stream>>>(d, a);
  PCK(cudaGetLastError());
  return 0;
}

// ===================================================================================================
// LSTM weight gradients on the tensor cores (replaces two cuBLAS GEMMs, a column reduction and the fp32 unpacking of
// X / Hp):   dWx += X^T dZ,   dWh += Hp^T dZ,   dbl += 1^T dZ      over the rows m = (t, replica) of one chunk.
// Same MN-major staging as fc_bwd_tc_kernel.  A = [X | Hp | 1] (dx + 64 + 1 columns = up to three M = 128 blocks),
// B = one 128-column half of dZ; D[block][128 gate columns] accumulates in TMEM (384 columns).  A CTA owns a
// contiguous tile range of the (unit, half, tile) list and flushes with atomics when (unit, half) changes.
// Hp is rebuilt from the bf16 activation store: Hp[t] = (1 - done[t]) * (t > 0 ? H[t-1] : h0).
#define WG_A_CHUNKS 40
#define WG_Z_CHUNKS 16
#define WG_STAGE ((WG_A_CHUNKS + WG_Z_CHUNKS) * FBT_SBO)
struct WGradTC {
  const float* dZ;             // [2A][M][256] fp32, or
  const __nv_bfloat16* dZb;    // [2A][M][256] bf16
  const float* X;              // [2A][M][dx] fp32, or
  const __nv_bfloat16* Xb;     // [2A][M][dx] bf16
  const float* Hp;             // [2A][M][64] fp32, or
  const __nv_bfloat16* Hb;     // [2A][T][rc][64] bf16 with h0 / done / T / rc / ld_state / r0
  const float* h0;
  const float* done;
  float* G;
  int64_t M, rc, ld_state, r0;
  int T, variant;
};

__global__ void __launch_bounds__(FBT_THREADS, 1)
wgrad_tc_kernel(const DDimsTC d, const WGradTC a) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(tc_smem + 2 * WG_STAGE);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 2);
  const uint32_t bar0 = smem_u32(sBar);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // never-written A chunks are read by the last M block: keep them finite
  for (int i = tid; i < 2 * WG_STAGE / 16; i += FBT_THREADS) reinterpret_cast<uint4*>(tc_smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *sTmem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  const uint32_t lbo = a.variant ? FBT_SBO : 128, sbo = a.variant ? 128 : FBT_SBO;
  const int dx = d.dx, ng = dx >> 3, n_xitems = FBT_ROWS * ng;
  const int nb = (dx + TC_H + 1 + 127) >> 7;        // M blocks
  const int64_t tpu = (a.M + FBT_ROWS - 1) / FBT_ROWS;
  const int64_t NT = tpu * 4 * d.A;
  const int64_t j0 = NT * blockIdx.x / gridDim.x, j1 = NT * (blockIdx.x + 1) / gridDim.x;
  uint32_t ph0 = 0, ph1 = 0;
  bool pend0 = false, pend1 = false, first = true;
  int cur_pu = -1;

  auto flush = [&](int pu) {
    if (pend0) { mbar_wait(bar0, ph0); ph0 ^= 1; pend0 = false; }
    if (pend1) { mbar_wait(bar0 + 8, ph1); ph1 ^= 1; pend1 = false; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int u = pu >> 1, nh = pu & 1;
    const int q = warp & 3, cq = warp >> 2;
    const int g0 = nh * 128 + cq * 32;
    for (int b = 0; b < nb; ++b) {
      const int ci = b * 128 + q * 32 + lane;
      float v[32];
      const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 128 + cq * 32);
      tmem_ld16(tb, v); tmem_ld16(tb + 16, v + 16);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* dst = nullptr;
      if (ci < dx) dst = a.G + d.off_wx + ((int64_t)u * dx + ci) * TC_N + g0;
      else if (ci < dx + TC_H) dst = a.G + d.off_wh + ((int64_t)u * TC_H + (ci - dx)) * TC_N + g0;
      else if (ci == dx + TC_H) dst = a.G + d.off_bl + (int64_t)u * TC_N + g0;
      if (dst) {
#pragma unroll
        for (int e = 0; e < 32; ++e) atomicAdd(dst + e, v[e]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  };

  for (int64_t j = j0; j < j1; ++j) {
    const int pu = (int)(j / tpu), u = pu >> 1, nh = pu & 1;
    const int64_t m0 = (j - (int64_t)pu * tpu) * FBT_ROWS;
    const int s = (int)((j - j0) & 1);
    if (pu != cur_pu) {
      if (cur_pu >= 0) flush(cur_pu);
      cur_pu = pu; first = true;
    }
    if (s == 0) { if (pend0) { mbar_wait(bar0, ph0); ph0 ^= 1; pend0 = false; } }
    else { if (pend1) { mbar_wait(bar0 + 8, ph1); ph1 ^= 1; pend1 = false; } }
    unsigned char* sA = tc_smem + (size_t)s * WG_STAGE;
    unsigned char* sZ = sA + (size_t)WG_A_CHUNKS * FBT_SBO;
    const int rows_valid = (a.M - m0) < FBT_ROWS ? (int)(a.M - m0) : FBT_ROWS;
    const int64_t rowbase = (int64_t)u * a.M + m0;
    if (j + 1 < j1) {      // pull the next tile towards L2 while this one is converted and multiplied
      const int pun = (int)((j + 1) / tpu), un = pun >> 1, nhn = pun & 1;
      const int64_t m0n = (j + 1 - (int64_t)pun * tpu) * FBT_ROWS;
      const int rvn = (a.M - m0n) < FBT_ROWS ? (int)(a.M - m0n) : FBT_ROWS;
      const int64_t rbn = (int64_t)un * a.M + m0n;
      if ((tid >> 2) < rvn) {
        if (a.dZb) { if ((tid & 3) < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dZb + (rbn + (tid >> 2)) * TC_N + nhn * 128 + (tid & 3) * 64)); }
        else asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dZ + (rbn + (tid >> 2)) * TC_N + nhn * 128 + (tid & 3) * 32));
      }
      if (a.Xb) {
        const char* px = reinterpret_cast<const char*>(a.Xb + rbn * dx);
        const int64_t nbytes = (int64_t)rvn * dx * 2;
        for (int64_t o = (int64_t)tid * 128; o < nbytes; o += FBT_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + o));
      }
      if (a.Hb && tid < rvn && m0n + tid >= a.rc) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Hb + (rbn + tid - a.rc) * TC_H));
    }
    // ---- B: this half of dZ (fp32 -> bf16), 8 gate columns per item ----
    {
      float4 z0[4], z1[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = r * FBT_THREADS + tid, row = i >> 4, g = i & 15;
        z0[r] = z1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows_valid) {
          if (a.dZb) {       // already bf16: carried bit-for-bit in z0
            const uint4 zb = __ldg(reinterpret_cast<const uint4*>(a.dZb + (rowbase + row) * TC_N + nh * 128 + g * 8));
            z0[r] = make_float4(__uint_as_float(zb.x), __uint_as_float(zb.y), __uint_as_float(zb.z), __uint_as_float(zb.w));
          } else {
            const float4* zp = reinterpret_cast<const float4*>(a.dZ + (rowbase + row) * TC_N + nh * 128 + g * 8);
            z0[r] = __ldg(zp); z1[r] = __ldg(zp + 1);
          }
        }
      }
      // ---- A: X, 8 columns per item; items are contiguous in global memory ----
      for (int ib = 0; ib < n_xitems; ib += 4 * FBT_THREADS) {
        uint4 xb[4];
        float4 x0[4], x1[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = ib + r * FBT_THREADS + tid;
          xb[r] = make_uint4(0, 0, 0, 0);
          x0[r] = x1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < rows_valid * ng) {
            if (a.Xb) xb[r] = __ldg(reinterpret_cast<const uint4*>(a.Xb + rowbase * dx) + i);
            else {
              const float4* xp = reinterpret_cast<const float4*>(a.X + rowbase * dx) + 2 * (int64_t)i;
              x0[r] = __ldg(xp); x1[r] = __ldg(xp + 1);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = ib + r * FBT_THREADS + tid;
          if (i < n_xitems) {
            uint4 o = xb[r];
            if (!a.Xb) {
              __align__(16) __nv_bfloat16 t[8] = {__float2bfloat16_rn(x0[r].x), __float2bfloat16_rn(x0[r].y),
                                                  __float2bfloat16_rn(x0[r].z), __float2bfloat16_rn(x0[r].w),
                                                  __float2bfloat16_rn(x1[r].x), __float2bfloat16_rn(x1[r].y),
                                                  __float2bfloat16_rn(x1[r].z), __float2bfloat16_rn(x1[r].w)};
              o = *reinterpret_cast<const uint4*>(t);
            }
            const int row = i / ng, cg = i - row * ng;
            *reinterpret_cast<uint4*>(sA + (size_t)cg * FBT_SBO + row * 16) = o;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = r * FBT_THREADS + tid, row = i >> 4, g = i & 15;
        __align__(16) __nv_bfloat16 t[8] = {__float2bfloat16_rn(z0[r].x), __float2bfloat16_rn(z0[r].y),
                                            __float2bfloat16_rn(z0[r].z), __float2bfloat16_rn(z0[r].w),
                                            __float2bfloat16_rn(z1[r].x), __float2bfloat16_rn(z1[r].y),
                                            __float2bfloat16_rn(z1[r].z), __float2bfloat16_rn(z1[r].w)};
        uint4 zo = *reinterpret_cast<const uint4*>(t);
        if (a.dZb) zo = make_uint4(__float_as_uint(z0[r].x), __float_as_uint(z0[r].y), __float_as_uint(z0[r].z), __float_as_uint(z0[r].w));
        *reinterpret_cast<uint4*>(sZ + (size_t)g * FBT_SBO + row * 16) = zo;
      }
    }
    // ---- A: Hp (8 hidden units per item) and the ones column ----
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = r * FBT_THREADS + tid, row = i >> 3, c = i & 7;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (row < rows_valid) {
        if (a.Hb) {
          const int64_t m = m0 + row;
          const int t = (int)(m / a.rc);
          if (a.done[t] == 0.f) {
            if (t > 0) o = __ldg(reinterpret_cast<const uint4*>(a.Hb + (rowbase + row - a.rc) * TC_H + c * 8));
            else {
              const float4* hp = reinterpret_cast<const float4*>(a.h0 + ((int64_t)u * a.ld_state + a.r0 + m) * TC_H + c * 8);
              const float4 p = __ldg(hp), qv = __ldg(hp + 1);
              __align__(16) __nv_bfloat16 tt[8] = {__float2bfloat16_rn(p.x), __float2bfloat16_rn(p.y), __float2bfloat16_rn(p.z),
                                                   __float2bfloat16_rn(p.w), __float2bfloat16_rn(qv.x), __float2bfloat16_rn(qv.y),
                                                   __float2bfloat16_rn(qv.z), __float2bfloat16_rn(qv.w)};
              o = *reinterpret_cast<const uint4*>(tt);
            }
          }
        } else {
          const float4* hp = reinterpret_cast<const float4*>(a.Hp + (rowbase + row) * TC_H + c * 8);
          const float4 p = __ldg(hp), qv = __ldg(hp + 1);
          __align__(16) __nv_bfloat16 tt[8] = {__float2bfloat16_rn(p.x), __float2bfloat16_rn(p.y), __float2bfloat16_rn(p.z),
                                               __float2bfloat16_rn(p.w), __float2bfloat16_rn(qv.x), __float2bfloat16_rn(qv.y),
                                               __float2bfloat16_rn(qv.z), __float2bfloat16_rn(qv.w)};
          o = *reinterpret_cast<const uint4*>(tt);
        }
      }
      *reinterpret_cast<uint4*>(sA + (size_t)(ng + c) * FBT_SBO + row * 16) = o;
    }
    if (tid < FBT_ROWS)
      *reinterpret_cast<uint4*>(sA + (size_t)(ng + 8) * FBT_SBO + tid * 16) = make_uint4(tid < rows_valid ? 0x3f80u : 0u, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aA = smem_u32(sA), aZ = smem_u32(sZ);
      for (int b = 0; b < nb; ++b)
        for (int ks = 0; ks < FBT_ROWS / 16; ++ks)
          umma_bf16(tmem + b * 128, make_desc(aA + b * 16 * FBT_SBO + ks * 256, lbo, sbo),
                    make_desc(aZ + ks * 256, lbo, sbo), idesc, (first && ks == 0) ? 0u : 1u);
      umma_commit(bar0 + 8 * s);
    }
    if (s == 0) pend0 = true; else pend1 = true;
    first = false;
  }
  if (cur_pu >= 0) flush(cur_pu);
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// All-bf16 variant (the training loop's): 64-row tiles, two smem stages + one register stage.  CTAs 2p and 2p+1 walk the
// same (unit, tile) range, one per 128-column half of dZ, so the second reader of an X / Hp tile finds it in L2.
#define WGA_ROWS 64
#define WGA_SBO (WGA_ROWS * 16 + 16)
#define WGA_STAGE ((WG_A_CHUNKS + WG_Z_CHUNKS) * WGA_SBO)
#define WGA_STAGES 2
#define WGA_MAXT 256
__global__ void __launch_bounds__(FBT_THREADS, 1)
wgrad_tc_async_kernel(const DDimsTC d, const WGradTC a) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(tc_smem + WGA_STAGES * WGA_STAGE);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + WGA_STAGES);
  float* sDone = reinterpret_cast<float*>(sTmem + 2);
  const uint32_t bar0 = smem_u32(sBar);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < WGA_STAGES; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // never-written A chunks are read by the last M block: keep them finite
  for (int i = tid; i < WGA_STAGES * WGA_STAGE / 16; i += FBT_THREADS) reinterpret_cast<uint4*>(tc_smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < a.T; i += FBT_THREADS) sDone[i] = a.done[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *sTmem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  const uint32_t lbo = a.variant ? WGA_SBO : 128, sbo = a.variant ? 128 : WGA_SBO;
  const int dx = d.dx, ng = dx >> 3, n_xitems = WGA_ROWS * ng;
  const uint32_t ng_magic = (1u << 20) / (uint32_t)ng + 1u;      // i / ng == (i * magic) >> 20 for i < 4096, ng <= 32
  const int nb = (dx + TC_H + 1 + 127) >> 7;        // M blocks
  const int64_t tpu = (a.M + WGA_ROWS - 1) / WGA_ROWS;
  const int64_t NT = tpu * 2 * d.A;
  const int npairs = gridDim.x >> 1, pb = blockIdx.x >> 1, nh = blockIdx.x & 1;
  const int64_t j0 = NT * pb / npairs, j1 = NT * (pb + 1) / npairs;
  uint32_t pend = 0, phase = 0;                      // bit s: a commit is outstanding on / the wait parity of stage s
  bool first = true;
  int cur_u = -1;
  auto wait_stage = [&](int s) {
    if (pend & (1u << s)) { mbar_wait(bar0 + 8 * s, (phase >> s) & 1u); phase ^= 1u << s; pend &= ~(1u << s); }
  };
  auto flush = [&](int u) {
    for (int s = 0; s < WGA_STAGES; ++s) wait_stage(s);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, cq = warp >> 2;
    const int g0 = nh * 128 + cq * 32;
    for (int b = 0; b < nb; ++b) {
      const int ci = b * 128 + q * 32 + lane;
      float v[32];
      const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 128 + cq * 32);
      tmem_ld16(tb, v); tmem_ld16(tb + 16, v + 16);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* dst = nullptr;
      if (ci < dx) dst = a.G + d.off_wx + ((int64_t)u * dx + ci) * TC_N + g0;
      else if (ci < dx + TC_H) dst = a.G + d.off_wh + ((int64_t)u * TC_H + (ci - dx)) * TC_N + g0;
      else if (ci == dx + TC_H) dst = a.G + d.off_bl + (int64_t)u * TC_N + g0;
      if (dst) {
#pragma unroll
        for (int e = 0; e < 32; ++e) atomicAdd(dst + e, v[e]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  };

  // running position of the NEXT tile to issue: unit, first row, and (t, row within t) of that row
  int iu = (int)(j0 / tpu);
  int64_t im0 = (j0 - (int64_t)iu * tpu) * WGA_ROWS;
  int64_t it_t = im0 / a.rc, it_rem = im0 - it_t * a.rc;
  // Register-staged pipeline: the 7 pieces (16 B each: 2 of dZ, <= 4 of X, 1 of Hp) of tile j + 2 are loaded into
  // registers while tile j is multiplied and tile j + 1 sits in the other smem stage; they are stored one iteration
  // later, when their latency has passed.  (cp.async was measured slower here: LDGSTS keeps its address registers
  // reserved until the copy completes, and the allocator's reuse of them stalls the warp for a memory latency.)
  uint4 pv[7];
  int prv = 0;                                   // rows_valid of the tile held in pv
  auto load_tile = [&]() {
    const int u = iu;
    const int64_t m0 = im0;
    const int rows_valid = (a.M - m0) < WGA_ROWS ? (int)(a.M - m0) : WGA_ROWS;
    const int64_t rowbase = (int64_t)u * a.M + m0;
    prv = rows_valid;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = r * FBT_THREADS + tid, row = i >> 4, g = i & 15;
      pv[r] = make_uint4(0, 0, 0, 0);
      if (row < rows_valid) pv[r] = __ldg(reinterpret_cast<const uint4*>(a.dZb + (rowbase + row) * TC_N + nh * 128 + g * 8));
    }
    const uint4* xsrc = reinterpret_cast<const uint4*>(a.Xb + rowbase * dx);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = k * FBT_THREADS + tid;
      pv[2 + k] = make_uint4(0, 0, 0, 0);
      if (i < rows_valid * ng) pv[2 + k] = __ldg(xsrc + i);
    }
    {
      const int hrow = tid >> 3, hc = tid & 7;
      pv[6] = make_uint4(0, 0, 0, 0);
      if (hrow < rows_valid) {
        int64_t t = it_t, rem = it_rem + hrow;
        while (rem >= a.rc) { rem -= a.rc; ++t; }
        if (sDone[t] == 0.f) {
          if (t > 0) pv[6] = __ldg(reinterpret_cast<const uint4*>(a.Hb + (rowbase + hrow - a.rc) * TC_H + hc * 8));
          else {       // first step of the rollout: Hp = h0 (fp32 state)
            const float4* hp = reinterpret_cast<const float4*>(a.h0 + ((int64_t)u * a.ld_state + a.r0 + rem) * TC_H + hc * 8);
            const float4 p = __ldg(hp), qv = __ldg(hp + 1);
            __align__(16) __nv_bfloat16 tt[8] = {__float2bfloat16_rn(p.x), __float2bfloat16_rn(p.y), __float2bfloat16_rn(p.z),
                                                 __float2bfloat16_rn(p.w), __float2bfloat16_rn(qv.x), __float2bfloat16_rn(qv.y),
                                                 __float2bfloat16_rn(qv.z), __float2bfloat16_rn(qv.w)};
            pv[6] = *reinterpret_cast<const uint4*>(tt);
          }
        }
      }
    }
    // advance the running position
    im0 += WGA_ROWS; it_rem += WGA_ROWS;
    while (it_rem >= a.rc) { it_rem -= a.rc; ++it_t; }
    if (im0 >= a.M) { ++iu; im0 = 0; it_t = 0; it_rem = 0; }
  };
  auto store_tile = [&](int s) {
    unsigned char* sA = tc_smem + (size_t)s * WGA_STAGE;
    unsigned char* sZ = sA + (size_t)WG_A_CHUNKS * WGA_SBO;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = r * FBT_THREADS + tid, row = i >> 4, g = i & 15;
      *reinterpret_cast<uint4*>(sZ + (size_t)g * WGA_SBO + row * 16) = pv[r];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = k * FBT_THREADS + tid;
      if (i < n_xitems) {
        const int row = (int)(((uint32_t)i * ng_magic) >> 20), cg = i - row * ng;
        *reinterpret_cast<uint4*>(sA + (size_t)cg * WGA_SBO + row * 16) = pv[2 + k];
      }
    }
    *reinterpret_cast<uint4*>(sA + (size_t)(ng + (tid & 7)) * WGA_SBO + (tid >> 3) * 16) = pv[6];
    if (tid < WGA_ROWS)
      *reinterpret_cast<uint4*>(sA + (size_t)(ng + 8) * WGA_SBO + tid * 16) = make_uint4(tid < prv ? 0x3f80u : 0u, 0, 0, 0);
  };

  if (j0 < j1) { load_tile(); store_tile(0); }
  if (j0 + 1 < j1) load_tile();
  for (int64_t j = j0; j < j1; ++j) {
    const int u = (int)(j / tpu);
    const int s = (int)((j - j0) & 1);
    if (u != cur_u) {
      if (cur_u >= 0) flush(cur_u);
      cur_u = u; first = true;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aA = smem_u32(tc_smem + (size_t)s * WGA_STAGE), aZ = aA + WG_A_CHUNKS * WGA_SBO;
      for (int b = 0; b < nb; ++b)
        for (int ks = 0; ks < WGA_ROWS / 16; ++ks)
          umma_bf16(tmem + b * 128, make_desc(aA + b * 16 * WGA_SBO + ks * 256, lbo, sbo),
                    make_desc(aZ + ks * 256, lbo, sbo), idesc, (first && ks == 0) ? 0u : 1u);
      umma_commit(bar0 + 8 * s);
    }
    pend |= 1u << s;
    first = false;
    if (j + 1 < j1) {
      wait_stage(s ^ 1);            // MMAs of tile j - 1 (issued one iteration ago) have drained the other stage
      store_tile(s ^ 1);            // tile j + 1, loaded one iteration ago
      if (j + 2 < j1) load_tile();
    }
  }
  if (cur_u >= 0) flush(cur_u);
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

extern "C" int tscl_wgrad_tc(tscl_handle* h, const float* dZ, const void* dz_bf16, const float* X, const void* x_bf16,
                             const float* Hp, const void* h_bf16, const float* h0, const float* done, int32_t T, int64_t rc,
                             int64_t ld_state, int64_t r0, float* grads, int32_t variant, void* stream) {
  if (!h || (!dZ && !dz_bf16) || (!X && !x_bf16) || (!Hp && !h_bf16) || !grads || T <= 0 || rc <= 0)
    return tsc_set_error("tscl_wgrad_tc: bad argument");
  if (h_bf16 && (!h0 || !done)) return tsc_set_error("tscl_wgrad_tc: h_bf16 needs h0 and done");
  PCK(cudaSetDevice(tscl_device_of(h)));
  const DDimsTC& d = *tscl_dims_of(h);
  if ((d.dx % 8) != 0 || d.dx / 8 + 9 > WG_A_CHUNKS) return tsc_set_error("tscl_wgrad_tc: dx must be a multiple of 8, <= 240");
  const size_t smem = 2 * (size_t)WG_STAGE + 32;
  const size_t smem_async = (size_t)WGA_STAGES * WGA_STAGE + 8 * WGA_STAGES + 8 + 4 * WGA_MAXT;
  static int attr_dev = -1;
  if (attr_dev != tscl_device_of(h)) {
    PCK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PCK(cudaFuncSetAttribute(wgrad_tc_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_async));
    attr_dev = tscl_device_of(h);
  }
  int n_sm = 0;
  PCK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, tscl_device_of(h)));
  const int64_t M = (int64_t)T * rc;
  const int64_t NT = ((M + FBT_ROWS - 1) / FBT_ROWS) * 4 * d.A;
  const int grid = (int)(NT < n_sm ? NT : n_sm);
  WGradTC a;
  a.dZ = dZ; a.dZb = (const __nv_bfloat16*)dz_bf16; a.X = X; a.Xb = (const __nv_bfloat16*)x_bf16; a.Hp = Hp; a.Hb = (const __nv_bfloat16*)h_bf16; a.h0 = h0;
  a.done = done; a.G = grads; a.M = M; a.rc = rc; a.ld_state = ld_state; a.r0 = r0; a.T = T; a.variant = variant;
  if (a.dZb && a.Xb && a.Hb && T <= WGA_MAXT) {
    const int64_t nt2 = ((M + WGA_ROWS - 1) / WGA_ROWS) * 2 * d.A;
    int g2 = (int)(nt2 < n_sm / 2 ? nt2 : n_sm / 2) * 2;      // CTA pairs
    if (g2 < 2) g2 = 2;
    wgrad_tc_async_kernel<<<g2, FBT_THREADS, smem_async, (cudaStream_t)stream>>>(d, a);
  }
  else wgrad_tc_kernel<<<grid, FBT_THREADS, smem, (cudaStream_t)stream>>>(d, a);
  PCK(cudaGetLastError());
  return 0;
}

// ===================================================================================================
// dX = dZ . Wx^T  (input gradient of the LSTM's x-projection, agents/utils.py:103-105 differentiated): a streaming GEMM
// [M x 256] . [256 x dx] per unit, 960 B of HBM traffic per row.  Warp-specialised persistent kernel:
//   warps 4-7  loaders : 16-byte cp.async of one swizzle atom (128 rows x 64 K, 16 KB) per stage, written in the
//                        SWIZZLE_128B K-major pattern (chunk ^ (row & 7)); 3 stages, completion signalled two stages late
//   warp  8    MMA     : 4 x tcgen05.mma (M = 128, N = dx, K = 16) per atom, B = the unit's Wx^T image resident in shared
//                        memory (fetched by one cp.async.bulk per unit), accumulators double-buffered in TMEM (2 x 256 cols)
//   warps 0-3  epilogue: tcgen05.ld -> bf16 -> padded row in shared memory -> one cp.async.bulk store per row (448 B),
//                        drained by the copy engine while the next tile is converted
// A CTA owns a contiguous range of the (unit, 128-row tile) list, so it changes unit at most twice.
#define DXK_THREADS 288
#define DXK_STAGES 3
#define DXK_STAGE_BYTES 16384
struct DxTC {
  const __nv_bfloat16* dZb;    // [2A][M][256]
  const __nv_bfloat16* Wxt;    // [2A][32][dx][8]  (tscl_pack_wxt)
  __nv_bfloat16* dXb;          // [2A][M][dx]
  int64_t M;
};
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {      // K-major, 128-byte swizzle, 8-row groups 1024 B apart
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(DXK_THREADS, 1)
dx_tc_kernel(const DDimsTC d, const DxTC a) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dx = d.dx, row_bytes = dx * 2, out_stride = row_bytes + 16;
  unsigned char* sStage = tc_smem;
  unsigned char* sB = sStage + DXK_STAGES * DXK_STAGE_BYTES;
  unsigned char* sOut = sB + (size_t)BW_KC * dx * 16;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(sOut + (size_t)128 * out_stride);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 12);
  const uint32_t bar_full = smem_u32(sBar), bar_empty = bar_full + 24, bar_accf = bar_full + 48, bar_acce = bar_full + 64,
                 bar_b = bar_full + 80, bar_d = bar_full + 88;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < DXK_STAGES; ++s) { mbar_init(bar_full + 8 * s, 128); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_accf + 8 * b, 1); mbar_init(bar_acce + 8 * b, 128); }
    mbar_init(bar_b, 1); mbar_init(bar_d, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *sTmem;
  const int64_t tpu = (a.M + 127) / 128;
  const int64_t NT = tpu * 2 * d.A;
  const int64_t j0 = NT * blockIdx.x / gridDim.x, j1 = NT * (blockIdx.x + 1) / gridDim.x;

  if (warp >= 4 && warp < 8) {
    // ---------------- loaders ----------------
    const int lt = tid - 128;
    const uint32_t aS = smem_u32(sStage);
    int64_t it = 0;
    for (int64_t j = j0; j < j1; ++j) {
      const int u = (int)(j / tpu);
      const int64_t m0 = (j - (int64_t)u * tpu) * 128;
      const __nv_bfloat16* src0 = a.dZb + (int64_t)u * a.M * TC_N;
      for (int at = 0; at < 4; ++at, ++it) {
        const int s = (int)(it % DXK_STAGES);
        const int64_t n = it / DXK_STAGES;
        if (n > 0) mbar_wait(bar_empty + 8 * s, (uint32_t)((n - 1) & 1));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int id = i * 128 + lt, rw = id >> 3, c = id & 7;
          const int64_t rr = m0 + rw < a.M ? m0 + rw : a.M - 1;
          cp_async16(aS + s * DXK_STAGE_BYTES + rw * 128 + ((c ^ (rw & 7)) << 4), src0 + rr * TC_N + at * 64 + c * 8);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (it >= 2) {
          asm volatile("cp.async.wait_group 2;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(bar_full + 8 * (int)((it - 2) % DXK_STAGES));
        }
      }
    }
    if (it >= 2) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_full + 8 * (int)((it - 2) % DXK_STAGES));
    }
    if (it >= 1) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_full + 8 * (int)((it - 1) % DXK_STAGES));
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(dx >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
      const uint32_t aS = smem_u32(sStage), aB = smem_u32(sB);
      const uint32_t b_lbo = (uint32_t)dx * 16;
      int64_t it = 0, tc = 0;
      int cur_u = -1;
      uint32_t bph = 0, dph = 0;
      for (int64_t j = j0; j < j1; ++j, ++tc) {
        const int u = (int)(j / tpu);
        if (u != cur_u) {
          if (cur_u >= 0) {                       // the MMAs in flight still read the previous unit's image
            umma_commit(bar_d);
            mbar_wait(bar_d, dph); dph ^= 1;
          }
          cur_u = u;
          const uint32_t bytes = (uint32_t)BW_KC * dx * 16;
          mbar_expect_tx(bar_b, bytes);
          bulk_g2s(aB, a.Wxt + (int64_t)u * BW_KC * dx * 8, bytes, bar_b);
          mbar_wait(bar_b, bph); bph ^= 1;
        }
        const int b = (int)(tc & 1);
        const int64_t nb = tc >> 1;
        if (nb > 0) mbar_wait(bar_acce + 8 * b, (uint32_t)((nb - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int at = 0; at < 4; ++at, ++it) {
          const int s = (int)(it % DXK_STAGES);
          mbar_wait(bar_full + 8 * s, (uint32_t)((it / DXK_STAGES) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + b * 256, make_desc_sw128(aS + s * DXK_STAGE_BYTES + k * 32),
                      make_desc(aB + (uint32_t)((at * 4 + k) * 2) * b_lbo, b_lbo, 128), idesc, (at | k) ? 1u : 0u);
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_accf + 8 * b);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: thread = row (TMEM lane) ----------------
    const int row = tid;
    unsigned char* my = sOut + (size_t)row * out_stride;
    const uint32_t my_s = smem_u32(my);
    int64_t tc = 0;
    for (int64_t j = j0; j < j1; ++j, ++tc) {
      const int u = (int)(j / tpu);
      const int64_t m0 = (j - (int64_t)u * tpu) * 128;
      const int b = (int)(tc & 1);
      mbar_wait(bar_accf + 8 * b, (uint32_t)((tc >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous store of this row has left shared memory
      const uint32_t tb = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * 256);
      for (int c0 = 0; c0 < dx; c0 += 32) {
        float v[32];
        tmem_ld16(tb + c0, v);
        if (c0 + 16 < dx) tmem_ld16(tb + c0 + 16, v + 16);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int nq = (c0 + 16 < dx) ? 4 : 2;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          if (qd < nq) {
            __align__(16) __nv_bfloat16 o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float2bfloat16_rn(v[qd * 8 + e]);
            *reinterpret_cast<uint4*>(my + (c0 + qd * 8) * 2) = *reinterpret_cast<const uint4*>(o);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_acce + 8 * b);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (m0 + row < a.M) {
        __nv_bfloat16* dst = a.dXb + ((int64_t)u * a.M + m0 + row) * dx;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(my_s), "r"(row_bytes) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

extern "C" int tscl_dx_tc(tscl_handle* h, const void* dz_bf16, const void* wxt_bf16, void* dx_bf16, int64_t M, void* stream) {
  if (!h || !dz_bf16 || !wxt_bf16 || !dx_bf16 || M <= 0) return tsc_set_error("tscl_dx_tc: bad argument");
  PCK(cudaSetDevice(tscl_device_of(h)));
  const DDimsTC& d = *tscl_dims_of(h);
  if (d.dx % 16 != 0 || d.dx > 256 || d.dx < 16) return tsc_set_error("tscl_dx_tc: dx must be a multiple of 16 in [16, 256]");
  const size_t smem = (size_t)DXK_STAGES * DXK_STAGE_BYTES + (size_t)BW_KC * d.dx * 16 + (size_t)128 * (d.dx * 2 + 16) + 12 * 8 + 16;
  if (smem > 232448) return tsc_set_error("tscl_dx_tc: shared memory budget exceeded");
  static int attr_dev = -1;
  if (attr_dev != tscl_device_of(h)) {
    PCK(cudaFuncSetAttribute(dx_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_dev = tscl_device_of(h);
  }
  int n_sm = 0;
  PCK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, tscl_device_of(h)));
  const int64_t NT = ((M + 127) / 128) * 2 * d.A;
  const int grid = (int)(NT < n_sm ? NT : n_sm);
  DxTC a;
  a.dZb = (const __nv_bfloat16*)dz_bf16; a.Wxt = (const __nv_bfloat16*)wxt_bf16; a.dXb = (__nv_bfloat16*)dx_bf16; a.M = M;
  dx_tc_kernel<<<grid, DXK_THREADS, smem, (cudaStream_t)stream>>>(d, a);
  PCK(cudaGetLastError());
  return 0;
}
