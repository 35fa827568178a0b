// qr.cu — tnb200_qr: reduced Householder QR (decompositions.qr, backends/numpy/decompositions.py:77-98;
// LAPACK geqrf + orgqr there, same reflector convention: beta = -sign(alpha) * |x|, so R's diagonal
// signs match numpy's unless non_negative_diagonal asks for the phase fix of :91-94).
//
// Working storage is column-contiguous so every reflector application is a coalesced
// dot + axpy over one column per CTA.  Per column: one "larfg" launch (norm + scale of the
// reflector) and one "larf" launch over the trailing columns; Q is then formed by applying the
// reflectors in reverse order to the first r columns of the identity.
#include "common.cuh"
#include "cplx.cuh"
#include <math.h>

namespace tnb {

int copy_strided(const tnb200_tensor_t* src, const tnb200_tensor_t* dst, int conj, cudaStream_t st);

__device__ inline double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;  // every thread gets the total
}
__device__ inline double block_sum_t(double v, double* red) { return block_sum(v, red); }
__device__ inline zd block_sum_t(zd v, double* red) { double a = block_sum(v.x, red); double b = block_sum(v.y, red); return zd{a, b}; }

// LAPACK xLARFG on column j: W[j:, j] -> (beta (real) on the diagonal, v[1:] below it), tau[j]
template <typename T>
__global__ void __launch_bounds__(256) qr_larfg_kernel(T* __restrict__ W, int64_t m, int j, T* __restrict__ tau) {
  __shared__ double red[8];
  T* col = W + (int64_t)j * m;
  double acc = 0.0;
  for (int64_t i = j + 1 + threadIdx.x; i < m; i += blockDim.x) acc += ab2(col[i]);
  const double sigma2 = block_sum(acc, red);
  const T alpha = col[j];
  if (sigma2 == 0.0 && im_(alpha) == 0.0) {
    if (threadIdx.x == 0) tau[j] = zero_<T>();
    return;
  }
  const double nrm = sqrt(ab2(alpha) + sigma2);
  const double beta = re_(alpha) >= 0.0 ? -nrm : nrm;
  const T scale = divz(one_<T>(), sub(alpha, mk(beta, 0.0, (T*)nullptr)));
  for (int64_t i = j + 1 + threadIdx.x; i < m; i += blockDim.x) col[i] = mul(col[i], scale);
  __syncthreads();
  if (threadIdx.x == 0) {
    tau[j] = mk((beta - re_(alpha)) / beta, -im_(alpha) / beta, (T*)nullptr);
    col[j] = mk(beta, 0.0, (T*)nullptr);
  }
}

// apply H_j (conj_tau = 0) or H_j^H (conj_tau = 1), H_j = I - tau v v^H, v = [1; V[j+1:, j]], to the
// columns c0 + blockIdx.x of X (rows j..m-1)
template <typename T>
__global__ void __launch_bounds__(128) qr_larf_kernel(const T* __restrict__ V, T* __restrict__ X, int64_t m, int j, int c0,
                                                      const T* __restrict__ tau, int conj_tau) {
  __shared__ double red[4];
  T t = tau[j];
  if (re_(t) == 0.0 && im_(t) == 0.0) return;
  if (conj_tau) t = cj(t);
  const T* v = V + (int64_t)j * m;
  T* x = X + (int64_t)(c0 + blockIdx.x) * m;
  T acc = threadIdx.x == 0 ? x[j] : zero_<T>();
  for (int64_t i = j + 1 + threadIdx.x; i < m; i += blockDim.x) fmacc(acc, cj(v[i]), x[i]);
  const T w = mul(block_sum_t(acc, red), t);
  if (threadIdx.x == 0) x[j] = sub(x[j], w);
  for (int64_t i = j + 1 + threadIdx.x; i < m; i += blockDim.x) x[i] = sub(x[i], mul(w, v[i]));
}

template <typename T>
__global__ void qr_init_q_kernel(T* Q, int64_t m, int r) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx < m * r) Q[idx] = (idx % m == idx / m) ? one_<T>() : zero_<T>();
}

// write q (m x r) and r (r x n) with the optional sign fix: phases = sign(diag(R)) (diag is real)
template <typename T>
__global__ void qr_writeout_kernel(const T* __restrict__ W, const T* __restrict__ Q, int64_t m, int64_t n, int r, int nonneg,
                                   T* __restrict__ q, int64_t q_s0, int64_t q_s1, T* __restrict__ rr, int64_t r_s0, int64_t r_s1) {
  const int64_t total_q = m * r, total_r = (int64_t)r * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total_q + total_r; idx += (int64_t)gridDim.x * blockDim.x) {
    if (idx < total_q) {
      int64_t c = idx / m, i = idx % m;
      double ph = 1.0;
      if (nonneg) { double d = re_(W[c * m + c]); ph = d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0); }
      q[i * q_s0 + c * q_s1] = mulr(Q[c * m + i], ph);
    } else {
      int64_t k = idx - total_q;
      int64_t c = k / r, i = k % r;       // R[i, c]
      T val = i <= c ? W[c * m + i] : zero_<T>();
      if (nonneg) { double d = re_(W[i * m + i]); val = mulr(val, d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0)); }
      rr[i * r_s0 + c * r_s1] = val;
    }
  }
}

template <typename T>
static int qr_real(const tnb200_tensor_t* a, const tnb200_tensor_t* q, const tnb200_tensor_t* r, int nonneg, cudaStream_t st) {
  const int64_t m = a->shape[0], n = a->shape[1];
  const int k = (int)(m < n ? m : n);
  if (m == 0 || n == 0) return 0;
  T *W = nullptr, *Q = nullptr, *tau = nullptr;
  int rc;
  if ((rc = ws_alloc((void**)&W, sizeof(T) * (size_t)m * n, st))) return rc;
  if ((rc = ws_alloc((void**)&Q, sizeof(T) * (size_t)m * k, st))) return rc;
  if ((rc = ws_alloc((void**)&tau, sizeof(T) * (size_t)k, st))) return rc;
  tnb200_tensor_t dst;
  dst.data = W; dst.dtype = a->dtype; dst.ndim = 2;
  dst.shape[0] = m; dst.shape[1] = n; dst.stride[0] = 1; dst.stride[1] = m;
  if ((rc = copy_strided(a, &dst, 0, st))) return rc;
  for (int j = 0; j < k; ++j) {
    qr_larfg_kernel<T><<<1, 256, 0, st>>>(W, m, j, tau);
    if (j + 1 < n) qr_larf_kernel<T><<<(unsigned)(n - j - 1), 128, 0, st>>>(W, W, m, j, j + 1, tau, 1);
  }
  qr_init_q_kernel<T><<<(unsigned)((m * k + 255) / 256), 256, 0, st>>>(Q, m, k);
  for (int j = k - 1; j >= 0; --j) qr_larf_kernel<T><<<(unsigned)(k - j), 128, 0, st>>>(W, Q, m, j, j, tau, 0);
  int64_t tot = m * k + (int64_t)k * n;
  int64_t blocks = (tot + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  qr_writeout_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(W, Q, m, n, k, nonneg, (T*)q->data, q->stride[0], q->stride[1],
                                                         (T*)r->data, r->stride[0], r->stride[1]);
  TNB_LAUNCH_CHECK();
  count_launch(3 * k + 2);
  ws_free(W, st); ws_free(Q, st); ws_free(tau, st);
  return 0;
}

}  // namespace tnb

using namespace tnb;

extern "C" int32_t tnb200_qr(const tnb200_tensor_t* a, const tnb200_tensor_t* q, const tnb200_tensor_t* r,
                             int32_t non_negative_diagonal, void* stream) {
  TNB_REQUIRE(valid_tensor(a) && valid_tensor(q) && valid_tensor(r), TNB200_ERR_INVALID, "qr: invalid tensor descriptor");
  TNB_REQUIRE(a->ndim == 2 && q->ndim == 2 && r->ndim == 2, TNB200_ERR_INVALID, "qr: expects matrices");
  const int64_t m = a->shape[0], n = a->shape[1], k = m < n ? m : n;
  TNB_REQUIRE(q->shape[0] == m && q->shape[1] == k && r->shape[0] == k && r->shape[1] == n, TNB200_ERR_INVALID,
              "qr: output shapes must be (m,k), (k,n) with k = min(m,n)");
  TNB_REQUIRE(q->dtype == a->dtype && r->dtype == a->dtype, TNB200_ERR_DTYPE, "qr: dtype mismatch");
  TNB_REQUIRE(m < (1LL << 31) && n < (1LL << 31), TNB200_ERR_UNSUPPORTED, "qr: matrix too large");
  set_kernel_name("qr_householder");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->dtype == TNB200_F64) return qr_real<double>(a, q, r, non_negative_diagonal, st);
  if (a->dtype == TNB200_C128) return qr_real<zd>(a, q, r, non_negative_diagonal, st);
  if (a->dtype == TNB200_F32 || a->dtype == TNB200_C64) {   // widen, factor, round back
    const bool cplx = a->dtype == TNB200_C64;
    const int wide = cplx ? TNB200_C128 : TNB200_F64;
    const size_t esz = cplx ? 16 : 8;
    void *da = nullptr, *dq = nullptr, *dr = nullptr;
    int rc;
    if ((rc = ws_alloc(&da, esz * (size_t)m * n, st))) return rc;
    if ((rc = ws_alloc(&dq, esz * (size_t)m * k, st))) return rc;
    if ((rc = ws_alloc(&dr, esz * (size_t)k * n, st))) return rc;
    auto mk2 = [&](void* p, int64_t d0, int64_t d1) {
      tnb200_tensor_t t; t.data = p; t.dtype = wide; t.ndim = 2; t.shape[0] = d0; t.shape[1] = d1; t.stride[0] = d1; t.stride[1] = 1; return t;
    };
    tnb200_tensor_t ta = mk2(da, m, n), tq = mk2(dq, m, k), tr = mk2(dr, k, n);
    if ((rc = copy_strided(a, &ta, 0, st))) return rc;
    rc = cplx ? qr_real<zd>(&ta, &tq, &tr, non_negative_diagonal, st) : qr_real<double>(&ta, &tq, &tr, non_negative_diagonal, st);
    if (rc == 0) rc = copy_strided(&tq, q, 0, st);
    if (rc == 0) rc = copy_strided(&tr, r, 0, st);
    ws_free(da, st); ws_free(dq, st); ws_free(dr, st);
    return rc;
  }
  set_error("qr: dtype %s is not supported (f32/f64/c64/c128)", dtype_name(a->dtype));
  return TNB200_ERR_UNSUPPORTED;
}
