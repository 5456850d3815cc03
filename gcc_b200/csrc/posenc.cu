// posenc.cu -- Laplacian positional features of every ego-net in a batch, on device.
//
// Replaces _add_undirected_graph_positional_embedding + eigen_decomposision
// (gcc/datasets/data_util.py:242-281): L = D^-1/2 A D^-1/2 (D = in_deg.clip(1)),
// k = min(n-2, pos_dim) largest-algebraic eigenvectors in ascending order,
// row-L2 normalised (sklearn normalize), float32, zero-padded to pos_dim;
// k <= 0 -> zeros.  The reference runs ARPACK (scipy eigsh, float64, random v0)
// per ego-net on a CPU worker: ~2.8 ms each, the dominant cost of its pipeline.
//
// Method (B200): one CTA per ego-net, the dense matrix G = L + 2I (SPD, spectrum
// in [1,3]) lives in shared memory, column-major.  One-sided (Hestenes) Jacobi
// orthogonalises the columns of G with plane rotations applied on the right;
// because G is symmetric positive definite the converged columns are
// lambda_j' * v_j, so eigenvectors are the normalised columns and no separate
// V matrix is stored (n <= 232 fits in 227 KB).  A warp owns one column pair per
// step of a round-robin tournament (n/2 disjoint pairs per round), dot products
// by shuffle reduction, rotations in registers.  Degenerate spectra (ego-nets are
// star-like) are handled exactly: Jacobi returns an orthonormal basis of every
// eigenspace.  Ego-nets are binned by size into three launches (shared memory
// 17 KB / 66 KB / 218 KB) through device-built work lists.
#include "common.cuh"

namespace gccb {

#define GCCB_EIG_NMAX 232
#define GCCB_EIG_MAXSWEEP 14
#define GCCB_EIG_TOL 1.0e-6f

// class 0: n <= 64, class 1: n <= 128, class 2: n <= NMAX, class 3: larger (unsupported)
__device__ __forceinline__ int eig_class(int n) {
  return n <= 64 ? 0 : n <= 128 ? 1 : n <= GCCB_EIG_NMAX ? 2 : 3;
}

// Work lists: worklist[c][i] = slot.  One CTA, deterministic order.  grid = 1, block = 256.
__global__ void __launch_bounds__(256)
posenc_classify_kernel(const int64_t* __restrict__ counters, const int32_t* __restrict__ node_off,
                       int B, int32_t* __restrict__ worklist, int32_t* __restrict__ counts,
                       int32_t* __restrict__ flags) {
  __shared__ int scan_scratch[33];
  const int tid = threadIdx.x;
  int base[4] = {0, 0, 0, 0};
  for (int s0 = 0; s0 < 2 * B; s0 += 256) {
    int slot = s0 + tid;
    int cls = -1;
    if (slot < 2 * B) {
      int view = slot / B;
      if (node_off[view * (B + 1) + B] >= 0) cls = eig_class((int)counters[(size_t)slot * 4]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int tot;
      int ex = block_scan_excl(cls == c ? 1 : 0, scan_scratch, &tot);
      if (cls == c) worklist[(size_t)c * 2 * B + base[c] + ex] = slot;
      base[c] += tot;
    }
  }
  if (tid < 4) counts[tid] = base[tid];
  if (tid == 0 && base[3] > 0) atomicOr(flags, (int)GCCB_FLAG_EIG_TOOBIG);
}

// zero rows of ego-nets the eigensolver cannot take (class 3) so no garbage reaches the encoder
__global__ void posenc_zero_big_kernel(const int32_t* __restrict__ worklist,
                                       const int32_t* __restrict__ counts, int B, int node_cap,
                                       const int32_t* __restrict__ node_off, int pos_dim,
                                       float* __restrict__ pos, float* __restrict__ eigvals) {
  if ((int)blockIdx.x >= counts[3]) return;
  const int slot = worklist[(size_t)3 * 2 * B + blockIdx.x];
  const int view = slot / B, g = slot - view * B;
  const int noff = node_off[view * (B + 1) + g], n = node_off[view * (B + 1) + g + 1] - noff;
  float* out = pos + ((size_t)view * node_cap + noff) * pos_dim;
  for (int i = threadIdx.x; i < n * pos_dim; i += blockDim.x) out[i] = 0.f;
  if (eigvals)
    for (int i = threadIdx.x; i < pos_dim; i += blockDim.x) eigvals[(size_t)slot * pos_dim + i] = 0.f;
}

// One-sided Jacobi eigensolver + feature write.  NR = ceil(nmax / 32) rows per lane.
template <int NR, int THREADS>
__global__ void __launch_bounds__(THREADS)
posenc_jacobi_kernel(const int32_t* __restrict__ worklist, const int32_t* __restrict__ counts,
                     int cls, int B, int node_cap, int edge_cap,
                     const int32_t* __restrict__ node_off, const int32_t* __restrict__ b_indptr,
                     const int32_t* __restrict__ b_indices, const int32_t* __restrict__ sub_deg,
                     int pos_dim, int normalize, float* __restrict__ pos,
                     float* __restrict__ eigvals, int32_t* __restrict__ flags) {
  GCCB_DYN_SMEM(float, smem);
  __shared__ int sel[32];
  __shared__ float sgn[32];
  if ((int)blockIdx.x >= counts[cls]) return;
  const int slot = worklist[(size_t)cls * 2 * B + blockIdx.x];
  const int view = slot / B, g = slot - view * B;
  const int noff = node_off[view * (B + 1) + g];
  const int n = node_off[view * (B + 1) + g + 1] - noff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  const int k = min(n - 2, pos_dim);
  float* out = pos + ((size_t)view * node_cap + noff) * pos_dim;
  if (k <= 0) {                                        // data_util.py:243-244
    for (int i = tid; i < n * pos_dim; i += THREADS) out[i] = 0.f;
    if (eigvals)
      for (int i = tid; i < pos_dim; i += THREADS) eigvals[(size_t)slot * pos_dim + i] = 0.f;
    return;
  }
  const int ld = n;
  float* G = smem;                 // [n][ld] column-major: G[col * ld + row]
  float* nrm = G + (size_t)n * ld; // [n] squared column norms
  float* dinv = nrm + n;           // [n] D^-1/2
  const int32_t* v_indptr = b_indptr + (size_t)view * (node_cap + 1);
  const int32_t* v_indices = b_indices + (size_t)view * edge_cap;
  const int32_t* v_deg = sub_deg + (size_t)view * node_cap;

  for (int i = tid; i < n * ld; i += THREADS) G[i] = 0.f;
  for (int i = tid; i < n; i += THREADS) {
    int d = v_deg[noff + i];
    dinv[i] = 1.0f / sqrtf((float)(d < 1 ? 1 : d));   // in_degrees().clip(1) ** -0.5
  }
  __syncthreads();
  for (int i = warp; i < n; i += NW) {                 // row i <- its in-neighbours j
    const int beg = v_indptr[noff + i], end = v_indptr[noff + i + 1];
    const float di = dinv[i];
    for (int e = beg + lane; e < end; e += 32) {
      int j = v_indices[e] - noff;
      atomicAdd(&G[(size_t)j * ld + i], di * dinv[j]);
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += THREADS) G[(size_t)i * ld + i] += 2.0f;
  __syncthreads();

  const int m = n + (n & 1);                            // tournament size (even)
  int sweep = 0;
  for (; sweep < GCCB_EIG_MAXSWEEP; ++sweep) {
    for (int p = warp; p < n; p += NW) {                // exact column norms once per sweep
      float s = 0.f;
#pragma unroll
      for (int jj = 0; jj < NR; ++jj) {
        int r = lane + 32 * jj;
        float x = r < n ? G[(size_t)p * ld + r] : 0.f;
        s = fmaf(x, x, s);
      }
      s = warp_sum(s);
      if (lane == 0) nrm[p] = s;
    }
    __syncthreads();
    int rotated = 0;
    for (int r = 0; r < m - 1; ++r) {
      for (int pi = warp; pi < m / 2; pi += NW) {
        int p, q;
        if (pi == 0) { p = m - 1; q = r; }
        else { p = (r + pi) % (m - 1); q = (r + m - 1 - pi) % (m - 1); }
        if (p >= n || q >= n) continue;                 // bye (odd n)
        if (p > q) { int t = p; p = q; q = t; }
        float gp[NR], gq[NR];
        float gam = 0.f;
#pragma unroll
        for (int jj = 0; jj < NR; ++jj) {
          int rr = lane + 32 * jj;
          gp[jj] = rr < n ? G[(size_t)p * ld + rr] : 0.f;
          gq[jj] = rr < n ? G[(size_t)q * ld + rr] : 0.f;
          gam = fmaf(gp[jj], gq[jj], gam);
        }
        // read the cached norms BEFORE the shuffle reduction: the shuffles are the
        // convergence point that orders these reads against lane 0's update below
        const float alpha = nrm[p], beta = nrm[q];
        gam = warp_sum(gam);
        if (fabsf(gam) > GCCB_EIG_TOL * sqrtf(alpha * beta)) {      // warp-uniform
          const float zeta = (beta - alpha) / (2.0f * gam);
          const float t = (zeta >= 0.f ? 1.0f : -1.0f) / (fabsf(zeta) + sqrtf(1.0f + zeta * zeta));
          const float c = 1.0f / sqrtf(1.0f + t * t);
          const float s = c * t;
#pragma unroll
          for (int jj = 0; jj < NR; ++jj) {
            int rr = lane + 32 * jj;
            if (rr < n) {
              G[(size_t)p * ld + rr] = c * gp[jj] - s * gq[jj];
              G[(size_t)q * ld + rr] = s * gp[jj] + c * gq[jj];
            }
          }
          if (lane == 0) { nrm[p] = alpha - t * gam; nrm[q] = beta + t * gam; }
          rotated = 1;
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(rotated)) break;
  }
  if (sweep == GCCB_EIG_MAXSWEEP && tid == 0) atomicOr(flags, (int)GCCB_FLAG_EIG_NOCONV);

  // eigenvalue of column j = ||g_j|| - 2; rank columns, keep the k largest, ascending
  float* mu = nrm;                                      // reuse: mu[j] = ||g_j||
  for (int p = warp; p < n; p += NW) {
    float s = 0.f;
#pragma unroll
    for (int jj = 0; jj < NR; ++jj) {
      int r = lane + 32 * jj;
      float x = r < n ? G[(size_t)p * ld + r] : 0.f;
      s = fmaf(x, x, s);
    }
    s = warp_sum(s);
    if (lane == 0) mu[p] = sqrtf(s);
  }
  __syncthreads();
  for (int j = tid; j < n; j += THREADS) {
    const float mj = mu[j];
    int rank = 0;
    for (int i = 0; i < n; ++i) {
      float mi = mu[i];
      rank += (mi > mj) || (mi == mj && i < j);
    }
    if (rank < k) sel[k - 1 - rank] = j;                // ascending: slot k-1 = largest
  }
  __syncthreads();
  // deterministic sign: the largest-|.| component (lowest row on ties) is positive
  for (int c = warp; c < k; c += NW) {
    const float* col = G + (size_t)sel[c] * ld;
    float best = -1.f; int brow = 0x7fffffff;
    for (int r = lane; r < n; r += 32) {
      float a = fabsf(col[r]);
      if (a > best) { best = a; brow = r; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int orow = __shfl_xor_sync(0xffffffffu, brow, o);
      if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
    }
    if (lane == 0) sgn[c] = col[brow] < 0.f ? -1.0f : 1.0f;
  }
  __syncthreads();
  if (eigvals) {
    // eigenvalues as Rayleigh quotients v^T L v against the ORIGINAL sparse matrix: the
    // column norms carry the accumulated rounding of ~n rotations per sweep (~1e-5), the
    // Rayleigh quotient is second-order accurate in the eigenvector error.
    for (int c = warp; c < pos_dim; c += NW) {
      float acc = 0.f;
      if (c < k) {
        const float* col = G + (size_t)sel[c] * ld;
        for (int i = lane; i < n; i += 32) {
          const int beg = v_indptr[noff + i], end = v_indptr[noff + i + 1];
          float rowacc = 0.f;
          for (int e = beg; e < end; ++e) {
            int j = v_indices[e] - noff;
            rowacc = fmaf(dinv[j], col[j], rowacc);
          }
          acc = fmaf(col[i] * dinv[i], rowacc, acc);
        }
        acc = warp_sum(acc);
        const float m2 = mu[sel[c]];
        acc = acc / (m2 * m2);
      }
      if (lane == 0) eigvals[(size_t)slot * pos_dim + c] = acc;
    }
  }
  // rows: lane c holds component c of node r (pos_dim <= 32)
  for (int r = warp; r < n; r += NW) {
    float u = 0.f;
    if (lane < k) {
      const int j = sel[lane];
      u = sgn[lane] * G[(size_t)j * ld + r] / mu[j];
    }
    if (normalize) {                                    // sklearn normalize(norm="l2")
      float ss = warp_sum(u * u);
      if (ss > 0.f) u = u / sqrtf(ss);
    }
    if (lane < pos_dim) out[(size_t)r * pos_dim + lane] = u;
  }
}

static size_t eig_smem_bytes(int nmax) { return ((size_t)nmax * nmax + 2 * (size_t)nmax) * sizeof(float); }

}  // namespace gccb

using namespace gccb;

// workspace: worklist[4][2B] ints | counts[4] ints
extern "C" size_t gccb_posenc_workspace(int32_t batch, int32_t node_cap) {
  (void)node_cap;
  return ((size_t)4 * 2 * batch + 4) * sizeof(int32_t);
}

extern "C" int gccb_posenc(const gccb_batch_t* batch, int32_t pos_dim, int32_t normalize,
                           float* pos, float* eigvals, void* workspace, size_t workspace_bytes,
                           gccb_stream_t stream) {
  if (!batch || !pos || !workspace || batch->batch <= 0 || pos_dim < 2 || pos_dim > 32) {
    set_last_error("gccb_posenc: bad argument (pos_dim must be in [2, 32])");
    return GCCB_ERR_BADARG;
  }
  const int B = batch->batch;
  if (workspace_bytes < gccb_posenc_workspace(B, batch->node_cap)) {
    set_last_error("gccb_posenc: workspace too small");
    return GCCB_ERR_CAPACITY;
  }
  int32_t* worklist = (int32_t*)workspace;
  int32_t* counts = worklist + (size_t)4 * 2 * B;
  GCCB_LAUNCH(posenc_classify_kernel, 1, 256, 0, stream, batch->counters, batch->node_off, B,
              worklist, counts, batch->flags);
  auto k2 = posenc_jacobi_kernel<8, 1024>;
  auto k1 = posenc_jacobi_kernel<4, 512>;
  auto k0 = posenc_jacobi_kernel<2, 256>;
  const size_t s2 = eig_smem_bytes(GCCB_EIG_NMAX), s1 = eig_smem_bytes(128), s0 = eig_smem_bytes(64);
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2);
  cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1);
  // largest matrices first: they are the tail of the step
  GCCB_LAUNCH(k2, 2 * B, 1024, s2, stream, worklist, counts, 2, B, batch->node_cap,
              batch->edge_cap, batch->node_off, batch->indptr, batch->indices, batch->sub_deg,
              pos_dim, normalize, pos, eigvals, batch->flags);
  GCCB_LAUNCH(k1, 2 * B, 512, s1, stream, worklist, counts, 1, B, batch->node_cap,
              batch->edge_cap, batch->node_off, batch->indptr, batch->indices, batch->sub_deg,
              pos_dim, normalize, pos, eigvals, batch->flags);
  GCCB_LAUNCH(k0, 2 * B, 256, s0, stream, worklist, counts, 0, B, batch->node_cap,
              batch->edge_cap, batch->node_off, batch->indptr, batch->indices, batch->sub_deg,
              pos_dim, normalize, pos, eigvals, batch->flags);
  GCCB_LAUNCH(posenc_zero_big_kernel, 2 * B, 256, 0, stream, worklist, counts, B,
              batch->node_cap, batch->node_off, pos_dim, pos, eigvals);
  return check_launch("gccb_posenc");
}
