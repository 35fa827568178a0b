// svd.cu — tnb200_svd: thin SVD by blocked one-sided (Hestenes) Jacobi, and the truncation
// count of decompositions.svd (backends/numpy/decompositions.py:21-74; LAPACK gesdd there).
//
// The matrix is copied once into column-contiguous working storage W (tall: rows >= cols; a wide
// input is handled through its transpose).  Columns are grouped in blocks of SB; a sweep visits
// every block pair in a round-robin tournament (nb-1 rounds of nb/2 disjoint pairs, all pairs of
// a round processed concurrently):
//   1. gram   : G = [W_I W_J]^T [W_I W_J]           (2SB x 2SB per pair, split over row chunks)
//   2. eig    : cyclic Jacobi eigen-decomposition of G in shared memory -> rotation R
//   3. update : [W_I W_J] <- [W_I W_J] R,  [V_I V_J] <- [V_I V_J] R
// On convergence (all column pairs orthogonal to tol) sigma_j = |w_j|, U = W / sigma, and the
// triplets are sorted in descending order by a rank-counting kernel.  R is orthogonal to working
// precision, so the method is backward stable regardless of how accurately G was formed.
#include "common.cuh"
#include "cplx.cuh"
#include <math.h>
#include <stdlib.h>
#include <vector>

namespace tnb {

int copy_strided(const tnb200_tensor_t* src, const tnb200_tensor_t* dst, int conj, cudaStream_t st);

// Block width SB (columns per block; a pair rotates PB = 2 SB columns) is a template parameter: 16 is the default,
// 32 an experiment (see svd_dispatch): every round streams W and V through HBM once, and doubling the block width
// halves the number of rounds per sweep, but the Gram eigenproblem grows to 64 x 64 (still one CTA).
template <int SB> struct Geo {
  static constexpr int PB = 2 * SB;
  static constexpr int RT = SB == 16 ? 64 : 32;     // rows per shared-memory tile of the gram / update kernels
};

// round-robin tournament on nb (even) players: pair p of round r
__device__ __forceinline__ void rr_pair(int nb, int r, int p, int& i, int& j) {
  const int m = nb - 1;
  if (p == 0) { i = m; j = r % m; }
  else { i = (r + p) % m; j = (r - p + m) % m; }
  if (i > j) { int t = i; i = j; j = t; }
}
template <int SB>
__device__ __forceinline__ int pair_col(int bi, int bj, int c) { return c < SB ? bi * SB + c : bj * SB + (c - SB); }

template <typename T, int SB>
__global__ void __launch_bounds__(256) svd_gram_kernel(const T* __restrict__ W, int64_t R, int nb, int round, T* __restrict__ G, int rsplit) {
  constexpr int PB = Geo<SB>::PB, RT = Geo<SB>::RT, TPT = PB / 16;   // 16 x 16 threads, TPT x TPT outputs each
  __shared__ T tile[PB][RT + 1];
  const int pair = blockIdx.x, chunk = blockIdx.y;
  int bi, bj;
  rr_pair(nb, round, pair, bi, bj);
  const int64_t rows_per = ((R + rsplit - 1) / rsplit + RT - 1) / RT * RT;
  const int64_t r0 = chunk * rows_per, r1 = min(R, r0 + rows_per);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  T acc[TPT][TPT];
#pragma unroll
  for (int a = 0; a < TPT; ++a)
#pragma unroll
    for (int b = 0; b < TPT; ++b) acc[a][b] = zero_<T>();
  for (int64_t rb = r0; rb < r1; rb += RT) {
    for (int idx = threadIdx.x; idx < PB * RT; idx += 256) {
      int c = idx / RT, rr = idx % RT;
      int64_t row = rb + rr;
      tile[c][rr] = row < r1 ? W[(int64_t)pair_col<SB>(bi, bj, c) * R + row] : zero_<T>();
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < RT; ++rr) {
      T av[TPT], bv[TPT];
#pragma unroll
      for (int a = 0; a < TPT; ++a) { av[a] = cj(tile[ty * TPT + a][rr]); bv[a] = tile[tx * TPT + a][rr]; }
#pragma unroll
      for (int a = 0; a < TPT; ++a)
#pragma unroll
        for (int b = 0; b < TPT; ++b) fmacc(acc[a][b], av[a], bv[b]);
    }
    __syncthreads();
  }
  T* g = G + (int64_t)pair * PB * PB;     // G = W^H W (Hermitian)
#pragma unroll
  for (int a = 0; a < TPT; ++a)
#pragma unroll
    for (int b = 0; b < TPT; ++b) atomic_add(&g[(ty * TPT + a) * PB + tx * TPT + b], acc[a][b]);
}

// Diagonalise the PB x PB Gram matrix of each pair; write the rotation, clear G for the next round,
// record the largest relative off-diagonal seen BEFORE rotating (sweep convergence measure).
// Dynamic shared memory: g[PB][PB+1], rm[PB][PB+1] (T), then cs[SB], sn[SB] (double), ph[SB] (T), pp[SB], qq[SB] (int).
template <typename T, int SB>
__global__ void __launch_bounds__(256) svd_eig_kernel(T* __restrict__ G, T* __restrict__ Rout, unsigned int* conv, double tol_inner, int max_inner) {
  constexpr int PB = Geo<SB>::PB, LD = PB + 1;
  extern __shared__ __align__(16) unsigned char eig_smem[];
  T* g = reinterpret_cast<T*>(eig_smem);
  T* rm = g + PB * LD;
  double* cs = reinterpret_cast<double*>(rm + PB * LD);
  double* sn = cs + SB;
  T* ph = reinterpret_cast<T*>(sn + SB);          // e^{-i phi} of the pivot (real case: its sign is folded into t instead)
  int* pp = reinterpret_cast<int*>(ph + SB);
  int* qq = pp + SB;
  __shared__ float red[8];
  __shared__ float offmax;
  const int pair = blockIdx.x, tid = threadIdx.x;
  T* gg = G + (int64_t)pair * PB * PB;
  for (int idx = tid; idx < PB * PB; idx += 256) {
    int i = idx / PB, j = idx % PB;
    g[i * LD + j] = gg[idx];
    rm[i * LD + j] = i == j ? one_<T>() : zero_<T>();
    gg[idx] = zero_<T>();
  }
  __syncthreads();
  for (int sweep = 0; sweep < max_inner; ++sweep) {
    // relative off-diagonal measure
    float loc = 0.f;
    for (int idx = tid; idx < PB * PB; idx += 256) {
      int i = idx / PB, j = idx % PB;
      if (i < j) {
        double d = re_(g[i * LD + i]) * re_(g[j * LD + j]);
        if (d > 0.0) { float v = (float)(ab2(g[i * LD + j]) / d); loc = fmaxf(loc, v); }     // squared; root taken once below
      }
    }
    for (int o = 16; o > 0; o >>= 1) loc = fmaxf(loc, __shfl_xor_sync(0xffffffffu, loc, o));
    if ((tid & 31) == 0) red[tid >> 5] = loc;
    __syncthreads();
    if (tid == 0) {
      float m = 0.f;
      for (int w = 0; w < 8; ++w) m = fmaxf(m, red[w]);
      m = sqrtf(m);
      offmax = m;
      if (sweep == 0) atomicMax(conv, __float_as_uint(m));
    }
    __syncthreads();
    if (offmax <= (float)tol_inner) break;
    for (int step = 0; step < PB - 1; ++step) {
      if (tid < SB) {
        const int m = PB - 1;
        int p, q;
        if (tid == 0) { p = m; q = step % m; } else { p = (step + tid) % m; q = (step - tid + m) % m; }
        if (p > q) { int t = p; p = q; q = t; }
        // Hermitian 2x2 [[a, g], [conj g, b]], g = |g| e^{i phi}: rotate (x_p, e^{-i phi} x_q) by the
        // real Jacobi angle of [[a, |g|], [|g|, b]]
        const T gpq = g[p * LD + q];
        const double app = re_(g[p * LD + p]), aqq = re_(g[q * LD + q]);
        const double mag = mag_(gpq);
        double c = 1.0, s = 0.0;
        T e = one_<T>();
        if (mag > 1e-300) {
          e = unit_conj_phase(gpq);
          // t = sign(tau) / (|tau| + sqrt(1 + tau^2)), tau = (aqq - app) / (2 mag), written with one sqrt, one
          // division and one rsqrt (this scalar chain is the latency of every Jacobi step)
          const double dd = aqq - app, m2 = 2.0 * mag;
          const double t = (dd >= 0.0 ? m2 : -m2) / (fabs(dd) + sqrt(fma(dd, dd, m2 * m2)));
          c = rsqrt(fma(t, t, 1.0));
          s = t * c;
        }
        cs[tid] = c; sn[tid] = s; ph[tid] = e; pp[tid] = p; qq[tid] = q;
      }
      __syncthreads();
      // G <- J^H G J with J = the SB disjoint rotations of this step: the 2x2 block (rows p_i,q_i x columns p_j,q_j)
      // of every (row pair, column pair) is touched by exactly one thread, so the column rotation
      //   x_p' = c x_p - s e x_q ,  x_q' = s x_p + c e x_q            (e = e^{-i phi})
      // and the row rotation  r_p' = c r_p - s conj(e) r_q ,  r_q' = s r_p + c conj(e) r_q  are applied back to back
      // in registers, in place (same arithmetic, in the same order, as two separate passes — one barrier less per step)
      for (int blk = tid; blk < SB * SB; blk += 256) {
        const int ki = blk / SB, kj = blk % SB;
        const int pi = pp[ki], qi = qq[ki], pj = pp[kj], qj = qq[kj];
        const double cjj = cs[kj], sjj = sn[kj], cii = cs[ki], sii = sn[ki];
        const T ej = ph[kj], eic = cj(ph[ki]);
        const T a = g[pi * LD + pj], b = mul(ej, g[pi * LD + qj]), c2 = g[qi * LD + pj], d = mul(ej, g[qi * LD + qj]);
        const T a1 = sub(mulr(a, cjj), mulr(b, sjj)), b1 = add(mulr(a, sjj), mulr(b, cjj));
        const T c1 = sub(mulr(c2, cjj), mulr(d, sjj)), d1 = add(mulr(c2, sjj), mulr(d, cjj));
        const T yc = mul(eic, c1), yd = mul(eic, d1);
        g[pi * LD + pj] = sub(mulr(a1, cii), mulr(yc, sii)); g[qi * LD + pj] = add(mulr(a1, sii), mulr(yc, cii));
        g[pi * LD + qj] = sub(mulr(b1, cii), mulr(yd, sii)); g[qi * LD + qj] = add(mulr(b1, sii), mulr(yd, cii));
      }
      // accumulated eigenvector matrix: column rotations only
      for (int idx = tid; idx < SB * PB; idx += 256) {
        int k = idx / PB, i = idx % PB;
        const double c = cs[k], s = sn[k];
        const T e = ph[k];
        int p = pp[k], q = qq[k];
        T x = rm[i * LD + p], y = mul(e, rm[i * LD + q]);
        rm[i * LD + p] = sub(mulr(x, c), mulr(y, s)); rm[i * LD + q] = add(mulr(x, s), mulr(y, c));
      }
      __syncthreads();
    }
  }
  T* ro = Rout + (int64_t)pair * PB * PB;
  for (int idx = tid; idx < PB * PB; idx += 256) ro[idx] = rm[(idx / PB) * LD + idx % PB];
}
template <typename T, int SB>
static size_t eig_smem_bytes() {
  constexpr int PB = Geo<SB>::PB;
  return 2 * sizeof(T) * PB * (PB + 1) + SB * (2 * sizeof(double) + sizeof(T) + 2 * sizeof(int)) + 16;
}

// X[:, pair columns] <- X[:, pair columns] * R   (X = W or V; column-contiguous with `rows` rows)
template <typename T, int SB>
__global__ void __launch_bounds__(256) svd_update_kernel(T* __restrict__ X, int64_t rows, int nb, int round, const T* __restrict__ Rm) {
  constexpr int PB = Geo<SB>::PB, RT = Geo<SB>::RT, NCG = 256 / RT, CPT = PB / NCG;   // CPT = 8 outputs per thread
  __shared__ T tile[PB][RT];     // read as tile[k][row]: consecutive threads -> consecutive rows (no padding needed)
  __shared__ T rs[PB][PB];       // read as broadcast
  const int pair = blockIdx.x;
  int bi, bj;
  rr_pair(nb, round, pair, bi, bj);
  const T* rg = Rm + (int64_t)pair * PB * PB;
  for (int idx = threadIdx.x; idx < PB * PB; idx += 256) rs[idx / PB][idx % PB] = rg[idx];
  const int rr = threadIdx.x & (RT - 1), cg = threadIdx.x / RT;   // RT rows x NCG column groups of CPT
  for (int64_t rb = (int64_t)blockIdx.y * RT; rb < rows; rb += (int64_t)gridDim.y * RT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < PB * RT; idx += 256) {
      int c = idx / RT, r2 = idx % RT;
      int64_t row = rb + r2;
      tile[c][r2] = row < rows ? X[(int64_t)pair_col<SB>(bi, bj, c) * rows + row] : zero_<T>();
    }
    __syncthreads();
    T out[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) out[c] = zero_<T>();
#pragma unroll 8
    for (int k = 0; k < PB; ++k) {
      T x = tile[k][rr];
#pragma unroll
      for (int c = 0; c < CPT; ++c) fmacc(out[c], x, rs[k][cg * CPT + c]);
    }
    int64_t row = rb + rr;
    if (row < rows) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) X[(int64_t)pair_col<SB>(bi, bj, cg * CPT + c) * rows + row] = out[c];
    }
  }
}

template <typename T>
__global__ void svd_colnorm_kernel(const T* __restrict__ W, int64_t R, int ncols, double* __restrict__ sig) {
  const int j = blockIdx.x;
  if (j >= ncols) return;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < R; i += blockDim.x) acc += ab2(W[(int64_t)j * R + i]);
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    sig[j] = sqrt(t);
  }
}
// descending rank by counting (stable: ties keep column order)
__global__ void svd_rank_kernel(const double* __restrict__ sig, int n, int* __restrict__ rank) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double sj = sig[j];
  int r = 0;
  for (int i = 0; i < n; ++i) { double si = sig[i]; r += (si > sj) || (si == sj && i < j); }
  rank[j] = r;
}
// scatter the sorted triplets into the caller's u (m x r), s (r), vh (r x n); `tall` = input had m >= n
template <typename T>
__global__ void svd_finalize_kernel(const T* __restrict__ W, const T* __restrict__ V, const double* __restrict__ sig,
                                    const int* __restrict__ rank, int64_t R, int Cn, int Cp, int r_out, int tall,
                                    T* __restrict__ u, int64_t u_s0, int64_t u_s1, double* __restrict__ s, int64_t s_s0,
                                    T* __restrict__ vh, int64_t v_s0, int64_t v_s1) {
  const int j = blockIdx.x;          // working column
  const int k = rank[j];
  if (k >= r_out) return;
  const double sg = sig[j];
  const double inv = sg > 0.0 ? 1.0 / sg : 0.0;
  if (threadIdx.x == 0) s[(int64_t)k * s_s0] = sg;
  // WORK = A (tall) or A^H (wide) = Wn S V^H with Wn = W / sigma.
  //   tall: u = Wn, vh = V^H            wide: A = V S Wn^H  ->  u = V, vh = Wn^H
  for (int64_t i = threadIdx.x; i < R; i += blockDim.x) {
    T val = mulr(W[(int64_t)j * R + i], inv);
    if (tall) u[i * u_s0 + (int64_t)k * u_s1] = val; else vh[(int64_t)k * v_s0 + i * v_s1] = cj(val);
  }
  for (int64_t i = threadIdx.x; i < Cn; i += blockDim.x) {
    T val = V[(int64_t)j * Cp + i];
    if (tall) vh[(int64_t)k * v_s0 + i * v_s1] = cj(val); else u[i * u_s0 + (int64_t)k * u_s1] = val;
  }
}
template <typename T>
__global__ void svd_eye_kernel(T* V, int Cp) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx < (int64_t)Cp * Cp) V[idx] = (idx / Cp == idx % Cp) ? one_<T>() : zero_<T>();
}

template <typename T, int SB>
static int svd_real(const tnb200_tensor_t* a, const tnb200_tensor_t* u, const tnb200_tensor_t* s, const tnb200_tensor_t* vh,
                    int32_t* info_dev, cudaStream_t st) {
  constexpr int PB = Geo<SB>::PB, RT = Geo<SB>::RT;
  const int64_t m = a->shape[0], n = a->shape[1];
  const bool tall = m >= n;
  const int64_t R = tall ? m : n;
  const int Cn = (int)(tall ? n : m);
  const int Cp = (Cn + PB - 1) / PB * PB;
  const int nb = Cp / SB, npairs = nb / 2, rounds = nb - 1;
  if (Cn == 0 || R == 0) return 0;
  T *W = nullptr, *V = nullptr, *G = nullptr, *Rm = nullptr;
  double* sig = nullptr;
  int* rank = nullptr;
  unsigned int* conv = nullptr;
  int rc;
  if ((rc = ws_alloc((void**)&W, sizeof(T) * (size_t)Cp * R, st))) return rc;
  if ((rc = ws_alloc((void**)&V, sizeof(T) * (size_t)Cp * Cp, st))) return rc;
  if ((rc = ws_alloc((void**)&G, sizeof(T) * (size_t)npairs * PB * PB, st))) return rc;
  if ((rc = ws_alloc((void**)&Rm, sizeof(T) * (size_t)npairs * PB * PB, st))) return rc;
  if ((rc = ws_alloc((void**)&sig, sizeof(double) * (size_t)Cp, st))) return rc;
  if ((rc = ws_alloc((void**)&rank, sizeof(int) * (size_t)Cp, st))) return rc;
  if ((rc = ws_alloc((void**)&conv, sizeof(unsigned int) * 64, st))) return rc;
  TNB_CHECK_CUDA(cudaMemsetAsync(W, 0, sizeof(T) * (size_t)Cp * R, st));
  TNB_CHECK_CUDA(cudaMemsetAsync(G, 0, sizeof(T) * (size_t)npairs * PB * PB, st));
  // W[j * R + i] = a[i, j] (tall) or a[j, i] (wide)
  tnb200_tensor_t src = *a, dst;
  if (!tall) { src.shape[0] = a->shape[1]; src.shape[1] = a->shape[0]; src.stride[0] = a->stride[1]; src.stride[1] = a->stride[0]; }
  dst.data = W; dst.dtype = a->dtype; dst.ndim = 2;
  dst.shape[0] = R; dst.shape[1] = Cn; dst.stride[0] = 1; dst.stride[1] = R;
  if ((rc = copy_strided(&src, &dst, tall ? 0 : 1, st))) return rc;   // wide: WORK = A^H (conjugated)
  svd_eye_kernel<T><<<(unsigned)(((int64_t)Cp * Cp + 255) / 256), 256, 0, st>>>(V, Cp);
  count_launch();

  const double eps = 2.220446049250313e-16;
  const double tol = 4.0 * sqrt((double)R) * eps;
  // inner (Gram) eigen-solver: tolerance / sweep cap (env knobs for experiments; the outer criterion is unchanged)
  const char* e_tol = getenv("TNB200_SVD_INNER_TOL");
  const char* e_sw = getenv("TNB200_SVD_INNER_SWEEPS");
  const double tol_inner = e_tol ? atof(e_tol) : 1e-15;
  const int max_inner = e_sw ? atoi(e_sw) : 10;
  int rsplit = (4 * num_sms() + npairs - 1) / npairs;
  int max_split = (int)((R + 4 * RT - 1) / (4 * RT));
  if (rsplit > max_split) rsplit = max_split;
  if (rsplit < 1) rsplit = 1;
  int usplit_w = (int)((R + RT - 1) / RT); if (usplit_w > rsplit * 4) usplit_w = rsplit * 4;
  int usplit_v = (Cp + RT - 1) / RT; if (usplit_v > rsplit * 4) usplit_v = rsplit * 4;
  const size_t eig_bytes = eig_smem_bytes<T, SB>();
  {
    static bool attr_done = false;      // per (T, SB) instantiation
    if (!attr_done) {
      TNB_CHECK_CUDA(cudaFuncSetAttribute(svd_eig_kernel<T, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig_bytes));
      attr_done = true;
    }
  }
  const int max_sweeps = 40;
  int sweeps = 0, converged = 0;
  unsigned int h_conv = 0;
  for (int sw = 0; sw < max_sweeps; ++sw) {
    TNB_CHECK_CUDA(cudaMemsetAsync(conv, 0, sizeof(unsigned int), st));
    for (int r = 0; r < rounds; ++r) {
      svd_gram_kernel<T, SB><<<dim3(npairs, rsplit), 256, 0, st>>>(W, R, nb, r, G, rsplit);
      svd_eig_kernel<T, SB><<<npairs, 256, eig_bytes, st>>>(G, Rm, conv, tol_inner, max_inner);
      svd_update_kernel<T, SB><<<dim3(npairs, usplit_w), 256, 0, st>>>(W, R, nb, r, Rm);
      svd_update_kernel<T, SB><<<dim3(npairs, usplit_v), 256, 0, st>>>(V, Cp, nb, r, Rm);
    }
    count_launch(4 * rounds);
    TNB_LAUNCH_CHECK();
    TNB_CHECK_CUDA(cudaMemcpyAsync(&h_conv, conv, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    TNB_CHECK_CUDA(cudaStreamSynchronize(st));
    ++sweeps;
    float off;
    memcpy(&off, &h_conv, 4);
    if ((double)off <= tol) { converged = 1; break; }
  }
  svd_colnorm_kernel<T><<<Cp, 256, 0, st>>>(W, R, Cp, sig);
  svd_rank_kernel<<<(Cp + 255) / 256, 256, 0, st>>>(sig, Cp, rank);
  svd_finalize_kernel<T><<<Cp, 256, 0, st>>>(W, V, sig, rank, R, Cn, Cp, Cn, tall ? 1 : 0, (T*)u->data, u->stride[0], u->stride[1],
                                             (double*)s->data, s->stride[0], (T*)vh->data, vh->stride[0], vh->stride[1]);
  count_launch(3);
  TNB_LAUNCH_CHECK();
  if (info_dev) {
    int32_t h[4] = {sweeps, converged, 0, 0};
    TNB_CHECK_CUDA(cudaMemcpyAsync(info_dev, h, sizeof(h), cudaMemcpyHostToDevice, st));
    TNB_CHECK_CUDA(cudaStreamSynchronize(st));
  }
  ws_free(W, st); ws_free(V, st); ws_free(G, st); ws_free(Rm, st); ws_free(sig, st); ws_free(rank, st); ws_free(conv, st);
  if (!converged) { set_error("svd: Jacobi did not converge in %d sweeps", max_sweeps); return TNB200_ERR_NOCONV; }
  return 0;
}

// decompositions.py:38-57 on the device, in the arithmetic type of `s` (sequential cumsum like numpy)
template <typename T>
__global__ void svd_trunc_kernel(const T* __restrict__ s, int64_t n, int64_t stride, int64_t max_sv, int use_err, double max_err,
                                 int relative, long long* keep) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long by_err = max_sv;
  if (use_err) {
    T eps = relative ? (T)((T)max_err * s[0]) : (T)max_err;
    T cum = T(0);
    long long cnt = 0;
    for (int64_t i = n - 1; i >= 0; --i) {
      T v = s[i * stride];
      cum = cum + v * v;
      if (sqrt(cum) > eps) ++cnt;
    }
    by_err = cnt;
  }
  *keep = max_sv < by_err ? max_sv : by_err;
}

}  // namespace tnb

using namespace tnb;

// Block width: 16.  The 32-wide variant (half the rounds per sweep, i.e. half the HBM traffic of the gram / update
// kernels) is kept for real dtypes behind TNB200_SVD_SB=32, but it is SLOWER on B200 — measured 2048^2: 0.57 s vs
// 0.37 s, 4096^2: 2.20 s vs 2.06 s, same sweep counts — because the 64 x 64 Gram eigenproblem (63 dependent Jacobi
// steps per inner sweep in one CTA) then dominates every round.
static int svd_dispatch(bool cplx, const tnb200_tensor_t* a, const tnb200_tensor_t* u, const tnb200_tensor_t* s, const tnb200_tensor_t* vh,
                        int32_t* info_dev, cudaStream_t st) {
  if (cplx) return svd_real<zd, 16>(a, u, s, vh, info_dev, st);
  const int64_t cn = a->shape[0] < a->shape[1] ? a->shape[0] : a->shape[1];
  int sb = 16;
  (void)cn;
  if (const char* e = getenv("TNB200_SVD_SB")) { const int v = atoi(e); if (v == 16 || v == 32) sb = v; }
  return sb == 32 ? svd_real<double, 32>(a, u, s, vh, info_dev, st) : svd_real<double, 16>(a, u, s, vh, info_dev, st);
}

extern "C" int32_t tnb200_svd(const tnb200_tensor_t* a, const tnb200_tensor_t* u, const tnb200_tensor_t* s, const tnb200_tensor_t* vh,
                              int32_t* info_dev, void* stream) {
  TNB_REQUIRE(valid_tensor(a) && valid_tensor(u) && valid_tensor(s) && valid_tensor(vh), TNB200_ERR_INVALID, "svd: invalid tensor descriptor");
  TNB_REQUIRE(a->ndim == 2 && u->ndim == 2 && vh->ndim == 2 && s->ndim == 1, TNB200_ERR_INVALID, "svd: expects matrix arguments");
  const int64_t m = a->shape[0], n = a->shape[1], r = m < n ? m : n;
  TNB_REQUIRE(u->shape[0] == m && u->shape[1] == r && vh->shape[0] == r && vh->shape[1] == n && s->shape[0] == r, TNB200_ERR_INVALID,
              "svd: output shapes must be (m,r), (r,), (r,n) with r = min(m,n)");
  TNB_REQUIRE(u->dtype == a->dtype && vh->dtype == a->dtype, TNB200_ERR_DTYPE, "svd: u/vh dtype must equal the input dtype");
  TNB_REQUIRE(m < (1LL << 31) && n < (1LL << 31), TNB200_ERR_UNSUPPORTED, "svd: matrix too large");
  cudaStream_t st = (cudaStream_t)stream;
  set_kernel_name("svd_block_jacobi");
  if (a->dtype == TNB200_F64) { TNB_REQUIRE(s->dtype == TNB200_F64, TNB200_ERR_DTYPE, "svd: s must be f64"); return svd_dispatch(false, a, u, s, vh, info_dev, st); }
  if (a->dtype == TNB200_C128) { TNB_REQUIRE(s->dtype == TNB200_F64, TNB200_ERR_DTYPE, "svd: s must be f64"); return svd_dispatch(true, a, u, s, vh, info_dev, st); }
  if (a->dtype == TNB200_F32 || a->dtype == TNB200_C64) {
    // single precision input: iterate in double (hundreds of accumulated plane rotations cost ~1e-5
    // relative accuracy in fp32, LAPACK's sgesdd delivers ~1e-6), then round the factors back.
    const bool cplx = a->dtype == TNB200_C64;
    TNB_REQUIRE(s->dtype == TNB200_F32, TNB200_ERR_DTYPE, "svd: s must be f32");
    const int wide_dt = cplx ? TNB200_C128 : TNB200_F64;
    const size_t esz = cplx ? 16 : 8;
    void *da = nullptr, *du = nullptr, *dv = nullptr;
    double* ds = nullptr;
    int rc;
    if ((rc = ws_alloc(&da, esz * (size_t)m * n, st))) return rc;
    if ((rc = ws_alloc(&du, esz * (size_t)m * r, st))) return rc;
    if ((rc = ws_alloc((void**)&ds, sizeof(double) * (size_t)r, st))) return rc;
    if ((rc = ws_alloc(&dv, esz * (size_t)r * n, st))) return rc;
    auto mk = [](void* p, int dt, int64_t d0, int64_t d1, int nd) {
      tnb200_tensor_t t; t.data = p; t.dtype = dt; t.ndim = nd;
      t.shape[0] = d0; t.shape[1] = d1; t.stride[0] = nd == 2 ? d1 : 1; t.stride[1] = 1; return t;
    };
    tnb200_tensor_t ta = mk(da, wide_dt, m, n, 2), tu = mk(du, wide_dt, m, r, 2), ts = mk(ds, TNB200_F64, r, 1, 1), tv = mk(dv, wide_dt, r, n, 2);
    if ((rc = copy_strided(a, &ta, 0, st))) return rc;
    rc = svd_dispatch(cplx, &ta, &tu, &ts, &tv, info_dev, st);
    if (rc == 0) rc = copy_strided(&tu, u, 0, st);
    if (rc == 0) rc = copy_strided(&ts, s, 0, st);
    if (rc == 0) rc = copy_strided(&tv, vh, 0, st);
    ws_free(da, st); ws_free(du, st); ws_free(ds, st); ws_free(dv, st);
    return rc;
  }
  set_error("svd: dtype %s is not supported (f32/f64/c64/c128)", dtype_name(a->dtype));
  return TNB200_ERR_UNSUPPORTED;
}

extern "C" int32_t tnb200_svd_truncation_count(const tnb200_tensor_t* s, int64_t max_singular_values, int32_t use_error,
                                               double max_truncation_error, int32_t relative, int64_t* keep_dev, void* stream) {
  TNB_REQUIRE(valid_tensor(s) && s->ndim == 1 && keep_dev, TNB200_ERR_INVALID, "svd_truncation_count: invalid arguments");
  const int64_t n = s->shape[0];
  int64_t max_sv = max_singular_values < 0 ? n : max_singular_values;
  cudaStream_t st = (cudaStream_t)stream;
  if (s->dtype == TNB200_F64)
    svd_trunc_kernel<double><<<1, 32, 0, st>>>((const double*)s->data, n, s->stride[0], max_sv, use_error, max_truncation_error, relative, (long long*)keep_dev);
  else if (s->dtype == TNB200_F32)
    svd_trunc_kernel<float><<<1, 32, 0, st>>>((const float*)s->data, n, s->stride[0], max_sv, use_error, max_truncation_error, relative, (long long*)keep_dev);
  else { set_error("svd_truncation_count: s must be f32/f64"); return TNB200_ERR_DTYPE; }
  TNB_LAUNCH_CHECK();
  count_launch();
  return 0;
}
