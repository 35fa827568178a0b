// tensordot_skinny.cu — CUDA-core kernels for the two degenerate GEMM shapes of a greedy MPS contraction
// path, where the tensor cores have nothing to chew on and the work is pure HBM streaming:
//
//  (1) skinny_outer_kernel : one side tiny (S <= 16 rows) AND a short contraction (K <= 32), the other
//      side huge (L up to 10^6 x batch) — the "ramp-up" steps, e.g. (4 x 4) . (4 x 131072).  One thread
//      per element of the long side: it streams its K inputs (coalesced across threads), multiplies with
//      the tiny operand held in shared memory and writes its S outputs.  Algorithmic bytes
//      (L*K + L*S) * sizeof are moved exactly once.
//  (2) skinny_dot_kernel   : both sides tiny (M*N <= 16) and an enormous contraction (K >= 4096) — the
//      closing step (2 x 262144) . (262144 x 2).  grid = (batch, K-splits); threads stride over k, keep
//      M*N partial sums in registers, block-reduce, atomically accumulate into an fp32/fp64 workspace that
//      a finalize kernel converts.  Arbitrary (multi-mode) strides on every operand: nothing is repacked.
#include "common.cuh"

namespace tnb {

constexpr int SK_MAXS = 16, SK_MAXK = 32;

template <typename T>
struct SkinnyParams {
  const T* Lp; const T* Sp; T* C;
  DevModes mB;     // batch modes: s0 = long operand, s1 = short operand, s2 = C
  DevModes mL;     // long free modes : s0 = long operand, s1 = C
  DevModes mS;     // short free modes: s0 = short operand, s1 = C
  DevModes mK;     // contracted modes: s0 = long operand, s1 = short operand
  int64_t L, batch;
  int S, K;
};

template <typename T, typename Acc, int VEC>
__global__ void __launch_bounds__(256) skinny_outer_kernel(const __grid_constant__ SkinnyParams<T> p) {
  __shared__ Acc sh[SK_MAXS][SK_MAXK];          // the tiny operand, one batch entry
  __shared__ long long kofs[SK_MAXK];           // long-operand offset of each k
  __shared__ long long sofs[SK_MAXS];           // C offset of each short index
  const int64_t bb = blockIdx.y;
  int64_t offLb, offSb, offCb;
  mode_offsets3(p.mB, bb, offLb, offSb, offCb);
  for (int idx = threadIdx.x; idx < p.S * p.K; idx += blockDim.x) {
    int s = idx / p.K, k = idx % p.K;
    int64_t os, oc, okl, oks;
    mode_offsets(p.mS, s, os, oc);
    mode_offsets(p.mK, k, okl, oks);
    sh[s][k] = to_acc(p.Sp[offSb + os + oks]);
  }
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) { int64_t a, b; mode_offsets(p.mK, k, a, b); kofs[k] = a; }
  for (int s = threadIdx.x; s < p.S; s += blockDim.x) { int64_t a, b; mode_offsets(p.mS, s, a, b); sofs[s] = b; }
  __syncthreads();
  // each thread owns VEC consecutive elements of the long side (VEC > 1 only when the innermost long
  // mode is unit-stride in both the operand and C and VEC divides its extent: 16-byte loads / stores)
  struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };
  const int64_t nvec = p.L / VEC;
  for (int64_t lv = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lv < nvec; lv += (int64_t)gridDim.x * blockDim.x) {
    int64_t ol, oc;
    mode_offsets(p.mL, lv * VEC, ol, oc);
    const T* src = p.Lp + offLb + ol;
    Acc acc[SK_MAXS][VEC];
#pragma unroll
    for (int s = 0; s < SK_MAXS; ++s)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[s][v] = acc_zero((Acc*)nullptr);
    // 4 independent loads in flight per thread before any of them is consumed (latency hiding)
    for (int k0 = 0; k0 < p.K; k0 += 4) {
      Pack x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) if (k0 + j < p.K) x[j] = *reinterpret_cast<const Pack*>(src + kofs[k0 + j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) if (k0 + j < p.K) {
#pragma unroll
        for (int s = 0; s < SK_MAXS; ++s) if (s < p.S) {
          const Acc w = sh[s][k0 + j];
#pragma unroll
          for (int v = 0; v < VEC; ++v) fma_acc(acc[s][v], to_acc(x[j].v[v]), w);
        }
      }
    }
    T* dst = p.C + offCb + oc;
#pragma unroll
    for (int s = 0; s < SK_MAXS; ++s) if (s < p.S) {
      Pack o;
#pragma unroll
      for (int v = 0; v < VEC; ++v) o.v[v] = FromAcc<T, Acc>::f(acc[s][v]);
      *reinterpret_cast<Pack*>(dst + sofs[s]) = o;
    }
  }
}

template <typename T>
struct DotParams {
  const T* A; const T* B; T* C;
  DevModes mB, mM, mN, mK;   // as in the generic kernel: mM (s0 A, s1 C), mN (s0 B, s1 C), mK (s0 A, s1 B)
  int64_t K, batch, kchunk;
  int M, N, inner;
  void* ws;                  // [batch][M][N] accumulators
};

constexpr int DOT_IB = 2048;   // inner block of the contraction whose operand offsets are tabulated in smem

template <typename T, typename Acc>
__global__ void __launch_bounds__(256) skinny_dot_kernel(const __grid_constant__ DotParams<T> p) {
  __shared__ long long aofs[16], bofs[16];
  __shared__ int ia[DOT_IB], ib[DOT_IB];         // offsets of the inner block (relative, fit in 32 bits)
  __shared__ Acc red[8][16];
  const int64_t bb = blockIdx.x;
  int64_t offAb, offBb, offCb;
  mode_offsets3(p.mB, bb, offAb, offBb, offCb);
  if (threadIdx.x < p.M) { int64_t a, c; mode_offsets(p.mM, threadIdx.x, a, c); aofs[threadIdx.x] = a; }
  if (threadIdx.x < p.N) { int64_t b, c; mode_offsets(p.mN, threadIdx.x, b, c); bofs[threadIdx.x] = b; }
  const int inner = p.inner;                     // product of the trailing K modes (<= DOT_IB), divides K
  for (int i = threadIdx.x; i < inner; i += blockDim.x) {
    int64_t oa, ob;
    mode_offsets(p.mK, i, oa, ob);               // i < inner only touches the trailing modes
    ia[i] = (int)oa; ib[i] = (int)ob;
  }
  __syncthreads();
  Acc acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = acc_zero((Acc*)nullptr);
  const int64_t nouter = p.K / inner;
  const int64_t o0 = (int64_t)blockIdx.y * p.kchunk;            // kchunk counts OUTER indices here
  const int64_t o1 = o0 + p.kchunk < nouter ? o0 + p.kchunk : nouter;
  for (int64_t o = o0; o < o1; ++o) {
    int64_t oa, ob;
    mode_offsets(p.mK, o * inner, oa, ob);       // offsets of the outer K modes (uniform across the CTA)
    const T* Ap = p.A + offAb + oa;
    const T* Bp = p.B + offBb + ob;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) {
      const int xa = ia[i], xb = ib[i];
      Acc a[4], b[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) a[m] = m < p.M ? to_acc(Ap[aofs[m] + xa]) : acc_zero((Acc*)nullptr);
#pragma unroll
      for (int n = 0; n < 4; ++n) b[n] = n < p.N ? to_acc(Bp[bofs[n] + xb]) : acc_zero((Acc*)nullptr);
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) fma_acc(acc[m * 4 + n], a[m], b[n]);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    Acc v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int m = threadIdx.x / 4, n = threadIdx.x % 4;
    if (m < p.M && n < p.N) {
      Acc v = acc_zero((Acc*)nullptr);
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      atomicAdd((Acc*)p.ws + (bb * p.M + m) * p.N + n, v);
    }
  }
}
// Fast path of the same reduction when both operands are packed [K][M] / [K][N] with M == N == MN: every
// thread streams 16-byte vectors (16/sizeof(T)/MN whole k-rows each) of A and B, four vector pairs in flight.
template <typename T, typename Acc, int MN>
__global__ void __launch_bounds__(256) skinny_dot_vec_kernel(const __grid_constant__ DotParams<T> p) {
  constexpr int VE = 16 / (int)sizeof(T);
  constexpr int R = VE / MN;                     // k-rows per vector
  __shared__ Acc red[8][MN * MN];
  const int64_t bb = blockIdx.x;
  int64_t offAb, offBb, offCb;
  mode_offsets3(p.mB, bb, offAb, offBb, offCb);
  const T* Ap = p.A + offAb;
  const T* Bp = p.B + offBb;
  struct alignas(16) Pack { T v[VE]; };
  Acc acc[MN][MN];
#pragma unroll
  for (int m = 0; m < MN; ++m)
#pragma unroll
    for (int n = 0; n < MN; ++n) acc[m][n] = acc_zero((Acc*)nullptr);
  const int64_t nvec = p.K * MN / VE;
  const int64_t v0 = (int64_t)blockIdx.y * p.kchunk;             // kchunk counts vectors here
  const int64_t v1 = v0 + p.kchunk < nvec ? v0 + p.kchunk : nvec;
  constexpr int U = 4;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += 256 * U) {
    Pack a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (v + u * 256 < v1) {
        *reinterpret_cast<uint4*>(&a[u]) = __ldg(reinterpret_cast<const uint4*>(Ap + (v + u * 256) * VE));
        *reinterpret_cast<uint4*>(&b[u]) = __ldg(reinterpret_cast<const uint4*>(Bp + (v + u * 256) * VE));
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (v + u * 256 < v1) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int m = 0; m < MN; ++m) {
            const Acc av = to_acc(a[u].v[r * MN + m]);
#pragma unroll
            for (int n = 0; n < MN; ++n) fma_acc(acc[m][n], av, to_acc(b[u].v[r * MN + n]));
          }
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int m = 0; m < MN; ++m)
#pragma unroll
    for (int n = 0; n < MN; ++n) {
      Acc x = acc[m][n];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) red[warp][m * MN + n] = x;
    }
  __syncthreads();
  if (threadIdx.x < MN * MN) {
    Acc x = acc_zero((Acc*)nullptr);
    for (int w = 0; w < 8; ++w) x += red[w][threadIdx.x];
    atomicAdd((Acc*)p.ws + bb * MN * MN + threadIdx.x, x);
  }
}

template <typename T, typename Acc>
__global__ void skinny_dot_finalize(const __grid_constant__ DotParams<T> p) {
  const int64_t total = p.batch * p.M * p.N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i % p.N, t = i / p.N, m = t % p.M, bb = t / p.M;
    int64_t oa, ob, oc, o1, ocm, o2, ocn;
    mode_offsets3(p.mB, bb, oa, ob, oc);
    mode_offsets(p.mM, m, o1, ocm);
    mode_offsets(p.mN, n, o2, ocn);
    p.C[oc + ocm + ocn] = FromAcc<T, Acc>::f(((const Acc*)p.ws)[i]);
  }
}

template <int DT>
static int run_outer(const void* Lp, const void* Sp, void* C, const ModeList& mB, const ModeList& mL, const ModeList& mS,
                     const ModeList& mK, cudaStream_t st) {
  using T = typename DType<DT>::T;
  using Acc = typename DType<DT>::Acc;
  SkinnyParams<T> p;
  p.Lp = (const T*)Lp; p.Sp = (const T*)Sp; p.C = (T*)C;
  if (!to_dev(mB, p.mB) || !to_dev(mL, p.mL) || !to_dev(mS, p.mS) || !to_dev(mK, p.mK)) return TNB200_ERR_UNSUPPORTED;
  p.L = mL.total(); p.batch = mB.total(); p.S = (int)mS.total(); p.K = (int)mK.total();
  if (p.batch > 65535) return TNB200_ERR_UNSUPPORTED;
  // vector width: 16 bytes when the innermost long mode is contiguous in the operand and in C and everything is aligned
  // <= 4 elements per thread: S * VEC accumulators must stay in registers (16 x 8 floats spilled)
  constexpr int VMAX = (16 / (int)sizeof(T)) > 4 ? 4 : (16 / (int)sizeof(T));
  bool vec = mL.n > 0 && mL.s0[mL.n - 1] == 1 && mL.s1[mL.n - 1] == 1 && mL.ext[mL.n - 1] % VMAX == 0 &&
             ((uintptr_t)Lp % 16 == 0) && ((uintptr_t)C % 16 == 0);
  auto mult = [&](int64_t x) { return x % VMAX == 0; };
  for (int i = 0; i + 1 < mL.n && vec; ++i) vec = mult(mL.s0[i]) && mult(mL.s1[i]);
  for (int i = 0; i < mK.n && vec; ++i) vec = mult(mK.s0[i]);
  for (int i = 0; i < mS.n && vec; ++i) vec = mult(mS.s1[i]);
  for (int i = 0; i < mB.n && vec; ++i) vec = mult(mB.s0[i]) && mult(mB.s2[i]);
  const int64_t work = vec ? p.L / VMAX : p.L;
  int64_t blocks = (work + 255) / 256;
  // few, fat CTAs: the per-CTA prologue (offset tables in shared memory) must be amortised over many
  // grid-stride iterations — about 4 CTAs per SM in total
  const int64_t cap = ((int64_t)num_sms() * 4 + p.batch - 1) / p.batch;
  if (blocks > cap) blocks = cap < 1 ? 1 : cap;
  if (vec) skinny_outer_kernel<T, Acc, VMAX><<<dim3((unsigned)blocks, (unsigned)p.batch), 256, 0, st>>>(p);
  else skinny_outer_kernel<T, Acc, 1><<<dim3((unsigned)blocks, (unsigned)p.batch), 256, 0, st>>>(p);
  TNB_LAUNCH_CHECK();
  count_launch();
  set_kernel_name("skinny_outer");
  return 0;
}
template <int DT>
static int run_dot(const void* A, const void* B, void* C, const ModeList& mB, const ModeList& mM, const ModeList& mN,
                   const ModeList& mK, cudaStream_t st) {
  using T = typename DType<DT>::T;
  using Acc = typename DType<DT>::Acc;
  DotParams<T> p;
  p.A = (const T*)A; p.B = (const T*)B; p.C = (T*)C;
  if (!to_dev(mB, p.mB) || !to_dev(mM, p.mM) || !to_dev(mN, p.mN) || !to_dev(mK, p.mK)) return TNB200_ERR_UNSUPPORTED;
  p.K = mK.total(); p.batch = mB.total(); p.M = (int)mM.total(); p.N = (int)mN.total();
  if (p.batch >= (1LL << 31)) return TNB200_ERR_UNSUPPORTED;
  // packed fast path: A = [K][M], B = [K][N], M == N in {1, 2, 4}, 16-byte aligned rows of vectors
  {
    constexpr int VE = 16 / (int)sizeof(T);
    const bool packed = mK.n == 1 && p.M == p.N && (p.M == 1 || p.M == 2 || p.M == 4) && p.M <= VE &&
                        mK.s0[0] == p.M && mK.s1[0] == p.N &&
                        (p.M == 1 || (mM.n == 1 && mN.n == 1 && mM.s0[0] == 1 && mN.s0[0] == 1)) &&
                        (p.K * p.M) % VE == 0 && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && p.batch <= 65535;
    bool al = packed;
    for (int i = 0; i < mB.n && al; ++i) al = mB.s0[i] % VE == 0 && mB.s1[i] % VE == 0;
    if (al) {
      const int64_t nvec = p.K * p.M / VE;
      int64_t want = ((int64_t)num_sms() * 8 + p.batch - 1) / p.batch;
      const int64_t maxsp = (nvec + 1023) / 1024;
      int64_t sp = want < maxsp ? want : maxsp; if (sp < 1) sp = 1; if (sp > 65535) sp = 65535;
      p.kchunk = (nvec + sp - 1) / sp;
      sp = (nvec + p.kchunk - 1) / p.kchunk;
      const size_t bytes = sizeof(Acc) * (size_t)(p.batch * p.M * p.N);
      int rc = ws_alloc(&p.ws, bytes, st);
      if (rc) return rc;
      TNB_CHECK_CUDA(cudaMemsetAsync(p.ws, 0, bytes, st));
      dim3 grid((unsigned)p.batch, (unsigned)sp);
      if (p.M == 1) skinny_dot_vec_kernel<T, Acc, 1><<<grid, 256, 0, st>>>(p);
      else if (p.M == 2) skinny_dot_vec_kernel<T, Acc, 2><<<grid, 256, 0, st>>>(p);
      else skinny_dot_vec_kernel<T, Acc, 4><<<grid, 256, 0, st>>>(p);
      const int64_t tot = p.batch * p.M * p.N;
      skinny_dot_finalize<T, Acc><<<(unsigned)((tot + 255) / 256 < 1024 ? (tot + 255) / 256 : 1024), 256, 0, st>>>(p);
      TNB_LAUNCH_CHECK();
      count_launch(2);
      set_kernel_name("skinny_dot");
      return ws_free(p.ws, st);
    }
  }
  // inner block = trailing K modes whose product stays <= DOT_IB (single-mode K: split it evenly)
  int64_t inner = 1;
  {
    int i = mK.n - 1;
    while (i >= 0 && inner * mK.ext[i] <= DOT_IB) { inner *= mK.ext[i]; --i; }
    if (inner == 1 && mK.n > 0) {            // innermost mode alone exceeds the table: carve a divisor out of it
      int64_t e = mK.ext[mK.n - 1], d = DOT_IB;
      while (d > 1 && e % d) --d;
      // represent K as (..., e/d, d): handled by pushing an extra mode into the device list below
      inner = d;
      if (d > 1) {
        DevModes& k = p.mK;
        if (k.n >= kDevModes) return TNB200_ERR_UNSUPPORTED;
        const int last = k.n - 1;
        k.ext[k.n] = d; k.s0[k.n] = k.s0[last]; k.s1[k.n] = k.s1[last]; k.s2[k.n] = 0;
        k.ext[last] = e / d; k.s0[last] *= d; k.s1[last] *= d;
        ++k.n;
      }
    }
  }
  p.inner = (int)inner;
  const int64_t nouter = p.K / inner;
  int64_t want = ((int64_t)num_sms() * 4 + p.batch - 1) / p.batch;
  int64_t sp = want < nouter ? want : nouter; if (sp < 1) sp = 1; if (sp > 65535) sp = 65535;
  p.kchunk = (nouter + sp - 1) / sp;
  sp = (nouter + p.kchunk - 1) / p.kchunk;
  const size_t bytes = sizeof(Acc) * (size_t)(p.batch * p.M * p.N);
  int rc = ws_alloc(&p.ws, bytes, st);
  if (rc) return rc;
  TNB_CHECK_CUDA(cudaMemsetAsync(p.ws, 0, bytes, st));
  skinny_dot_kernel<T, Acc><<<dim3((unsigned)p.batch, (unsigned)sp), 256, 0, st>>>(p);
  const int64_t tot = p.batch * p.M * p.N;
  skinny_dot_finalize<T, Acc><<<(unsigned)((tot + 255) / 256 < 1024 ? (tot + 255) / 256 : 1024), 256, 0, st>>>(p);
  TNB_LAUNCH_CHECK();
  count_launch(2);
  set_kernel_name("skinny_dot");
  return ws_free(p.ws, st);
}

// Entry used by the planner.  Mode lists follow tensordot.cu's conventions:
//   mB: s0 A, s1 B, s2 C;  mM: s0 A, s1 C;  mN: s0 B, s1 C;  mK: s0 A, s1 B.
// Returns TNB200_ERR_UNSUPPORTED when the shape is not one of the two skinny families.
int tensordot_skinny(int dt, const void* A, const void* B, void* C, const ModeList& mB, const ModeList& mM,
                     const ModeList& mN, const ModeList& mK, cudaStream_t st) {
  if (dt != TNB200_F64 && dt != TNB200_F32 && dt != TNB200_F16 && dt != TNB200_BF16) return TNB200_ERR_UNSUPPORTED;
  const int64_t M = mM.total(), N = mN.total(), K = mK.total();
  // (2) both tiny, long contraction
  if (M <= 4 && N <= 4 && K >= 4096) {
    switch (dt) {
      case TNB200_F64: return run_dot<TNB200_F64>(A, B, C, mB, mM, mN, mK, st);
      case TNB200_F32: return run_dot<TNB200_F32>(A, B, C, mB, mM, mN, mK, st);
      case TNB200_F16: return run_dot<TNB200_F16>(A, B, C, mB, mM, mN, mK, st);
      default: return run_dot<TNB200_BF16>(A, B, C, mB, mM, mN, mK, st);
    }
  }
  // (1) one side tiny, short contraction, other side long
  if (K <= SK_MAXK && ((M <= SK_MAXS && N >= 1024) || (N <= SK_MAXS && M >= 1024))) {
    const bool a_short = M <= SK_MAXS && N >= 1024;
    ModeList bat, lon, sho, kk;
    for (int i = 0; i < mB.n; ++i) bat.push(mB.ext[i], a_short ? mB.s1[i] : mB.s0[i], a_short ? mB.s0[i] : mB.s1[i], mB.s2[i]);
    const ModeList& L_ = a_short ? mN : mM;
    const ModeList& S_ = a_short ? mM : mN;
    for (int i = 0; i < L_.n; ++i) lon.push(L_.ext[i], L_.s0[i], L_.s1[i]);
    for (int i = 0; i < S_.n; ++i) sho.push(S_.ext[i], S_.s0[i], S_.s1[i]);
    for (int i = 0; i < mK.n; ++i) kk.push(mK.ext[i], a_short ? mK.s1[i] : mK.s0[i], a_short ? mK.s0[i] : mK.s1[i]);
    const void* Lp = a_short ? B : A;
    const void* Sp = a_short ? A : B;
    switch (dt) {
      case TNB200_F64: return run_outer<TNB200_F64>(Lp, Sp, C, bat, lon, sho, kk, st);
      case TNB200_F32: return run_outer<TNB200_F32>(Lp, Sp, C, bat, lon, sho, kk, st);
      case TNB200_F16: return run_outer<TNB200_F16>(Lp, Sp, C, bat, lon, sho, kk, st);
      default: return run_outer<TNB200_BF16>(Lp, Sp, C, bat, lon, sho, kk, st);
    }
  }
  return TNB200_ERR_UNSUPPORTED;
}

}  // namespace tnb
