// blocksparse_maps.cu — tnb200_blocksparse_maps: the int64 element maps of a block-sparse matrix view, built ON THE DEVICE.
//
// Reference: block_sparse/blocksparse_utils.py:330-634 (`_find_diagonal_sparse_blocks`, `_find_transposed_diagonal_sparse_
// blocks`, `reduce_charges`) — numpy unique / intersect / fancy indexing on the host, recomputed per call unless the opt-in
// cache (caching.py:22-88) is on.  Here it is pure integer work on the GPU, bit-identical to the host construction
// (tests/test_gpu_blocksparse.py compares with the golden maps of the real reference):
//
//   data vector   = the elements of the dense tensor (stored leg order, row-major) whose total signed charge is zero,
//                   in ascending flat position.  The stored legs are cut into a LEFT and a RIGHT group; element e is
//                   the j-th partner of left state l: r = bucket[ -q_l ][ j ]   (right states bucketed by charge, ascending)
//   matrix view   = rows: legs order[0..partition), columns: the rest.  Sector of charge q holds the rows with fused
//                   charge q (ascending row index) x the columns with charge -q (ascending column index), row-major.
//   map[ sect_off[q] + rowrank(R) * ncols[q] + colrank(C) ] = e
//
// Kernels: fuse (mixed-radix decode -> charge bin of every state of a leg group), rank (position of a state among the
// states of equal charge: one CTA per charge bin, ballot prefix sums), scan (first element of every left state), bucket,
// element (binary search e -> (l, j), decode, scatter).  The small per-charge tables (counts, offsets) are charge-
// degeneracy arithmetic done by the caller on the host (a few dozen integers: histogram convolutions of the legs).
#include "common.cuh"

namespace tnb {

constexpr int BM_MAXLEGS = TNB200_MAX_NDIM;

struct LegGroup {                      // an ordered group of stored legs forming one product space
  int n;
  int leg[BM_MAXLEGS];                 // stored leg ids, most significant first
  long long dim[BM_MAXLEGS];
  long long coff[BM_MAXLEGS];          // offset of the leg's signed charges in the charge table
};

// bin of a charge: U(1): q + shift (shift = sum of max |charge| over all legs);  Z_N: q mod N
__device__ __forceinline__ int bm_bin(long long q, long long shift, long long mod) {
  if (mod > 0) { long long r = q % mod; if (r < 0) r += mod; return (int)r; }
  return (int)(q + shift);
}

__global__ void bm_fuse_kernel(LegGroup g, const long long* __restrict__ charges, long long N, long long shift, long long mod,
                               int* __restrict__ bin) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= N) return;
  long long rem = s, q = 0;
#pragma unroll 1
  for (int i = g.n - 1; i >= 0; --i) {
    const long long d = rem % g.dim[i];
    rem /= g.dim[i];
    q += charges[g.coff[i] + d];
  }
  bin[s] = bm_bin(q, shift, mod);
}

// rank[s] = number of states s' < s with bin[s'] == bin[s]; cnt[b] = number of states in bin b.  One CTA per bin.
__global__ void __launch_bounds__(1024) bm_rank_kernel(const int* __restrict__ bin, long long N, int* __restrict__ rank,
                                                       long long* __restrict__ cnt) {
  __shared__ int wsum[32];
  __shared__ int running_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) running_s = 0;
  __syncthreads();
  for (long long base = 0; base < N; base += 1024) {
    const long long i = base + tid;
    const bool f = i < N && bin[i] == b;
    const unsigned m = __ballot_sync(0xffffffffu, f);
    const int pre = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wsum[warp] = __popc(m);
    __syncthreads();
    int off = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { const int v = wsum[w]; if (w < warp) off += v; tot += v; }
    const int run = running_s;
    if (f) rank[i] = run + off + pre;
    __syncthreads();
    if (tid == 0) running_s = run + tot;
    __syncthreads();
  }
  if (tid == 0) cnt[b] = running_s;
}

// first[l] = sum over l' < l of cnt_right[ partner bin of l' ]   (exclusive scan, one CTA; first[NL] = nnz)
__global__ void __launch_bounds__(1024) bm_first_kernel(const int* __restrict__ bin_left, long long NL, const long long* __restrict__ cnt_right,
                                                        long long shift, long long mod, long long* __restrict__ first) {
  __shared__ long long wsum[32];
  __shared__ long long running_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) running_s = 0;
  __syncthreads();
  for (long long base = 0; base < NL; base += 1024) {
    const long long i = base + tid;
    long long v = 0;
    if (i < NL) {
      const int b = bin_left[i];
      const long long pb = mod > 0 ? (mod - b) % mod : 2 * shift - b;      // bin of the charge -q
      v = cnt_right[pb];
    }
    long long x = v;                                                       // inclusive warp scan
    for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    long long off = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { const long long t = wsum[w]; if (w < warp) off += t; tot += t; }
    const long long run = running_s;
    if (i < NL) first[i] = run + off + x - v;
    __syncthreads();
    if (tid == 0) running_s = run + tot;
    __syncthreads();
  }
  if (tid == 0) first[NL] = running_s;
}

// bucket[start[bin[r]] + rank[r]] = r
__global__ void bm_bucket_kernel(const int* __restrict__ bin, const int* __restrict__ rank, long long N,
                                 const long long* __restrict__ start, long long* __restrict__ bucket) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r < N) bucket[start[bin[r]] + rank[r]] = r;
}

struct ElemParams {
  LegGroup left, right;                       // the two stored halves
  long long row_mul[BM_MAXLEGS], col_mul[BM_MAXLEGS];   // per STORED leg: multiplier in the row / column index (0 if absent)
  int is_row[BM_MAXLEGS];
  long long coff[BM_MAXLEGS];                 // per stored leg: offset of its charges
  long long NL, NR, nnz, shift, mod;
};

__global__ void bm_element_kernel(ElemParams p, const long long* __restrict__ charges, const int* __restrict__ bin_left,
                                  const long long* __restrict__ first, const long long* __restrict__ start_right,
                                  const long long* __restrict__ bucket, const int* __restrict__ row_rank, const int* __restrict__ col_rank,
                                  const long long* __restrict__ sect_off, const long long* __restrict__ ncols,
                                  long long* __restrict__ map) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= p.nnz) return;
  // l = last left state with first[l] <= e
  long long lo = 0, hi = p.NL;                // first[NL] = nnz > e
  while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (first[mid] <= e) lo = mid; else hi = mid; }
  const long long l = lo, j = e - first[l];
  const int bl = bin_left[l];
  const long long pb = p.mod > 0 ? (p.mod - bl) % p.mod : 2 * p.shift - bl;
  const long long r = bucket[start_right[pb] + j];
  long long R = 0, C = 0, rq = 0;
  long long rem = l;
#pragma unroll 1
  for (int i = p.left.n - 1; i >= 0; --i) {
    const long long d = rem % p.left.dim[i];
    rem /= p.left.dim[i];
    const int t = p.left.leg[i];
    R += d * p.row_mul[t]; C += d * p.col_mul[t];
    if (p.is_row[t]) rq += charges[p.coff[t] + d];
  }
  rem = r;
#pragma unroll 1
  for (int i = p.right.n - 1; i >= 0; --i) {
    const long long d = rem % p.right.dim[i];
    rem /= p.right.dim[i];
    const int t = p.right.leg[i];
    R += d * p.row_mul[t]; C += d * p.col_mul[t];
    if (p.is_row[t]) rq += charges[p.coff[t] + d];
  }
  const int qb = bm_bin(rq, p.shift, p.mod);
  map[sect_off[qb] + (long long)row_rank[R] * ncols[qb] + col_rank[C]] = e;
}

}  // namespace tnb

using namespace tnb;

extern "C" int32_t tnb200_blocksparse_maps(int32_t nlegs, const int64_t* dims, const int64_t* charges_dev, const int64_t* leg_off,
                                           const int32_t* order, int32_t partition, int32_t split, int64_t modulus, int64_t shift,
                                           int32_t nbins, const int64_t* tables_dev, int64_t nnz, int64_t* map_dev, void* stream) {
  TNB_REQUIRE(nlegs >= 1 && nlegs <= BM_MAXLEGS && partition >= 0 && partition <= nlegs && split >= 0 && split <= nlegs, TNB200_ERR_INVALID,
              "blocksparse_maps: bad leg counts");
  TNB_REQUIRE(dims && charges_dev && leg_off && order && tables_dev && (map_dev || nnz == 0) && nbins >= 1, TNB200_ERR_INVALID,
              "blocksparse_maps: null pointer");
  if (nnz == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // tables_dev (int64, uploaded by the caller): [start_right (nbins)] [sect_off (nbins)] [ncols (nbins)]
  const long long* start_right = (const long long*)tables_dev;
  const long long* sect_off = start_right + nbins;
  const long long* ncols = sect_off + nbins;
  auto group = [&](const int* legs, int n) {
    LegGroup g; g.n = n;
    for (int i = 0; i < BM_MAXLEGS; ++i) { g.leg[i] = 0; g.dim[i] = 1; g.coff[i] = 0; }
    for (int i = 0; i < n; ++i) { g.leg[i] = legs[i]; g.dim[i] = dims[legs[i]]; g.coff[i] = leg_off[legs[i]]; }
    return g;
  };
  auto total = [&](const LegGroup& g) { long long t = 1; for (int i = 0; i < g.n; ++i) t *= g.dim[i]; return t; };
  int stored[BM_MAXLEGS];
  for (int i = 0; i < nlegs; ++i) stored[i] = i;
  const LegGroup gl = group(stored, split), gr = group(stored + split, nlegs - split);
  const LegGroup grow = group(order, partition), gcol = group(order + partition, nlegs - partition);
  const long long NL = total(gl), NR = total(gr), NRo = total(grow), NCo = total(gcol);
  TNB_REQUIRE(NL < (1LL << 31) && NR < (1LL << 31) && NRo < (1LL << 31) && NCo < (1LL << 31), TNB200_ERR_UNSUPPORTED,
              "blocksparse_maps: a leg group has more than 2^31 states");
  int *bin_l = nullptr, *bin_r = nullptr, *bin_ro = nullptr, *bin_co = nullptr, *rank_r = nullptr, *rank_ro = nullptr, *rank_co = nullptr;
  long long *cnt = nullptr, *first = nullptr, *bucket = nullptr;
  int rc;
  if ((rc = ws_alloc((void**)&bin_l, sizeof(int) * (size_t)NL, st))) return rc;
  if ((rc = ws_alloc((void**)&bin_r, sizeof(int) * (size_t)NR, st))) return rc;
  if ((rc = ws_alloc((void**)&bin_ro, sizeof(int) * (size_t)NRo, st))) return rc;
  if ((rc = ws_alloc((void**)&bin_co, sizeof(int) * (size_t)NCo, st))) return rc;
  if ((rc = ws_alloc((void**)&rank_r, sizeof(int) * (size_t)NR, st))) return rc;
  if ((rc = ws_alloc((void**)&rank_ro, sizeof(int) * (size_t)NRo, st))) return rc;
  if ((rc = ws_alloc((void**)&rank_co, sizeof(int) * (size_t)NCo, st))) return rc;
  if ((rc = ws_alloc((void**)&cnt, sizeof(long long) * (size_t)nbins * 3, st))) return rc;
  if ((rc = ws_alloc((void**)&first, sizeof(long long) * (size_t)(NL + 1), st))) return rc;
  if ((rc = ws_alloc((void**)&bucket, sizeof(long long) * (size_t)NR, st))) return rc;
  const long long* ch = (const long long*)charges_dev;
  auto blocks = [](long long n) { return (unsigned)((n + 255) / 256); };
  bm_fuse_kernel<<<blocks(NL), 256, 0, st>>>(gl, ch, NL, shift, modulus, bin_l);
  bm_fuse_kernel<<<blocks(NR), 256, 0, st>>>(gr, ch, NR, shift, modulus, bin_r);
  bm_fuse_kernel<<<blocks(NRo), 256, 0, st>>>(grow, ch, NRo, shift, modulus, bin_ro);
  bm_fuse_kernel<<<blocks(NCo), 256, 0, st>>>(gcol, ch, NCo, shift, modulus, bin_co);
  bm_rank_kernel<<<nbins, 1024, 0, st>>>(bin_r, NR, rank_r, cnt);
  bm_rank_kernel<<<nbins, 1024, 0, st>>>(bin_ro, NRo, rank_ro, cnt + nbins);
  bm_rank_kernel<<<nbins, 1024, 0, st>>>(bin_co, NCo, rank_co, cnt + 2 * nbins);
  bm_first_kernel<<<1, 1024, 0, st>>>(bin_l, NL, cnt, shift, modulus, first);
  bm_bucket_kernel<<<blocks(NR), 256, 0, st>>>(bin_r, rank_r, NR, start_right, bucket);
  ElemParams p;
  p.left = gl; p.right = gr; p.NL = NL; p.NR = NR; p.nnz = nnz; p.shift = shift; p.mod = modulus;
  for (int t = 0; t < BM_MAXLEGS; ++t) { p.row_mul[t] = 0; p.col_mul[t] = 0; p.is_row[t] = 0; p.coff[t] = 0; }
  for (int t = 0; t < nlegs; ++t) p.coff[t] = leg_off[t];
  { long long m = 1; for (int i = partition - 1; i >= 0; --i) { p.row_mul[order[i]] = m; p.is_row[order[i]] = 1; m *= dims[order[i]]; } }
  { long long m = 1; for (int i = nlegs - 1; i >= partition; --i) { p.col_mul[order[i]] = m; m *= dims[order[i]]; } }
  bm_element_kernel<<<blocks(nnz), 256, 0, st>>>(p, ch, bin_l, first, start_right, bucket, rank_ro, rank_co, sect_off, ncols, (long long*)map_dev);
  TNB_LAUNCH_CHECK();
  count_launch(10);
  set_kernel_name("blocksparse_maps");
  ws_free(bin_l, st); ws_free(bin_r, st); ws_free(bin_ro, st); ws_free(bin_co, st); ws_free(rank_r, st); ws_free(rank_ro, st);
  ws_free(rank_co, st); ws_free(cnt, st); ws_free(first, st); ws_free(bucket, st);
  return 0;
}
