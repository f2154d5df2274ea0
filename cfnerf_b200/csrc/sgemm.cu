// fp32 CUDA-core GEMM with fused epilogues — the arithmetic of the CFN_PREC_FP32 "check" mode (1e-5 parity bar)
// and of the round-1 training path (forward with saved activations, dgrad, split-K wgrad).
// C(m,n) = epi( [C(m,n) +] sum_k A(m,k) B(k,n) + bias[n] ), arbitrary element strides on A and B.
// 128x128x16 block tile, 256 threads, 8x8 register micro-tile, operands staged k-major in shared memory.
#include "common.cuh"

namespace cfn {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g, int64_t k_per_split) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * k_per_split;
  const int64_t k_end = min(g.K, k_begin + k_per_split);
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8 (m) x 8 (n)

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool a_kfast = (g.a_cs == 1);   // K contiguous in A (row-major activations)
  const bool b_kfast = (g.b_rs == 1);   // K contiguous in B (nn.Linear weight used as B(k,n) = W[n,k])

  for (int64_t kt = k_begin; kt < k_end; kt += BK) {
#pragma unroll
    for (int e = 0; e < (BM * BK) / 256; ++e) {
      const int idx = tid + e * 256;
      int m, k;
      if (a_kfast) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      const int64_t gm = m0 + m, gk = kt + k;
      As[k][m] = (gm < g.M && gk < k_end) ? __ldg(g.A + gm * g.a_rs + gk * g.a_cs) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < (BN * BK) / 256; ++e) {
      const int idx = tid + e * 256;
      int n, k;
      if (b_kfast) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
      const int gn = n0 + n;
      const int64_t gk = kt + k;
      Bs[k][n] = (gn < g.N && gk < k_end) ? __ldg(g.B + gk * g.b_rs + (int64_t)gn * g.b_cs) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4));
      if (gn >= g.N) continue;
      float* cp = g.C + gm * g.c_rs + gn;
      float v = acc[i][j];
      if (g.split_k > 1) {
        if (g.partials) g.partials[(int64_t)blockIdx.z * g.M * g.N + gm * g.N + gn] = v;   // deterministic mode: dense slab
        else atomicAdd(cp, v);
        continue;
      }
      if (g.bias) v += g.bias[gn];
      if (g.accumulate) v += *cp;
      if (g.epilogue == EPI_RELU) v = fmaxf(v, 0.f);
      else if (g.epilogue == EPI_TANH_MASK) { if (g.aux[gn] != 0.f) v = tanhf(v); }
      else if (g.epilogue == EPI_RELU_MASK_MUL) { if (!(g.aux[gm * g.aux_rs + gn] > 0.f)) v = 0.f; }
      *cp = v;
    }
  }
}

int launch_sgemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M == 0 || g.N == 0) return CFN_OK;
  CFN_CHECK_ARG(g.M > 0 && g.N > 0 && g.K >= 0, "sgemm: bad shape");
  int split = g.split_k > 1 ? g.split_k : 1;
  int64_t kps = (g.K + split - 1) / split;
  kps = ((kps + BK - 1) / BK) * BK;
  if (kps == 0) kps = BK;
  dim3 grid((unsigned)((g.M + BM - 1) / BM), (unsigned)((g.N + BN - 1) / BN), (unsigned)split);
  CFN_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "sgemm: grid too large");
  if (split > 1 && g.partials)
    CFN_CHECK_ARG((int64_t)split * g.M * g.N <= g.partials_floats, "sgemm: deterministic split-K scratch too small");
  sgemm_kernel<<<grid, 256, 0, s>>>(g, kps);
  CFN_LAUNCH_CHECK();
  if (split > 1 && g.partials) return reduce_split_partials(g.partials, split, g.M, g.N, g.C, g.c_rs, nullptr, nullptr, s);
  return CFN_OK;
}

// ---- second pass of the deterministic split-K: add the slabs in split order --------------------------------------
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int split, int64_t M, int N, float* __restrict__ C,
                                       int64_t c_rs, const float* __restrict__ rs_part, float* __restrict__ rs_out) {
  const int64_t total = M * (int64_t)N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total + (rs_out ? M : 0);
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i < total) {
      float acc = 0.f;
      for (int z = 0; z < split; ++z) acc += partials[(int64_t)z * total + i];
      C[(i / N) * c_rs + (i % N)] = acc;
    } else {
      const int64_t m = i - total;
      float acc = 0.f;
      for (int z = 0; z < split; ++z) acc += rs_part[(int64_t)z * M + m];
      rs_out[m] = acc;
    }
  }
}

int reduce_split_partials(const float* partials, int split, int64_t M, int N, float* C, int64_t c_rs,
                          const float* rowsum_partials, float* rowsum_out, cudaStream_t s) {
  const int64_t n = M * (int64_t)N + (rowsum_out ? M : 0);
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  reduce_partials_kernel<<<(unsigned)blocks, 256, 0, s>>>(partials, split, M, N, C, c_rs, rowsum_partials, rowsum_out);
  CFN_LAUNCH_CHECK();
  return CFN_OK;
}

}  // namespace cfn
