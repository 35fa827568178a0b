"""TEST INFRASTRUCTURE — numpy restatement of the reference's U(1) block-sparse contraction (CPU oracle /
cpu_baseline only; the product never imports this).

Follows tensornetwork/block_sparse/blocksparse_utils.py:330-425 (`_find_diagonal_sparse_blocks`: the
data vector holds the charge-allowed elements in row-major order, so every row of the matrix view whose
fused charge is a block charge stores its non-zeros contiguously) and
tensornetwork/block_sparse/blocksparsetensor.py:1031-1108 (`tensordot`: per-sector matmul, zero-initialised
result).  Pinned by tests/golden/blocksparse.npz (generated from the real reference).
"""
import numpy as np


def fuse(charges, flows):
  """fused U(1) charge of every multi-index (row-major); flow True = outflowing = dual charge (charge.py:62-75)"""
  out = np.zeros(1, dtype=np.int64)
  for q, f in zip(charges, flows):
    s = -1 if f else 1
    out = (out[:, None] + s * np.asarray(q, dtype=np.int64)[None, :]).ravel()
  return out


def num_nonzero(charges, flows):
  return int(np.count_nonzero(fuse(charges, flows) == 0))


def diagonal_blocks(charges, flows, partition):
  """blocksparse_utils.py:375-425: (block charges, {q: data positions (rows x cols)}, {q: (rows, cols)})"""
  rq = fuse(charges[:partition], flows[:partition])
  cq = fuse(charges[partition:], [not f for f in flows[partition:]])
  ucol, col_degen = np.unique(cq, return_counts=True)
  idx = np.minimum(np.searchsorted(ucol, rq), len(ucol) - 1)
  valid = ucol[idx] == rq
  row_nnz = np.where(valid, col_degen[idx], 0)
  starts = np.concatenate([[0], np.cumsum(row_nnz[:-1])]).astype(np.int64)
  maps, dims = {}, {}
  for q in np.intersect1d(np.unique(rq), ucol):
    rows = np.nonzero(rq == q)[0]
    cd = int(col_degen[np.searchsorted(ucol, q)])
    maps[int(q)] = (starts[rows][:, None] + np.arange(cd, dtype=np.int64)[None, :]).ravel()
    dims[int(q)] = (len(rows), cd)
  return maps, dims


def tensordot_trailing(data_a, charges_a, flows_a, data_b, charges_b, flows_b, n):
  """contract the last n legs of a with the last n legs of b (blocksparsetensor.py:1031-1108 with both
  operands viewed as (free | contracted)); returns (data_c, charges_c, flows_c)."""
  pa, pb = len(charges_a) - n, len(charges_b) - n
  ma, da = diagonal_blocks(charges_a, flows_a, pa)
  mb, db = diagonal_blocks(charges_b, flows_b, pb)
  charges_c = list(charges_a[:pa]) + list(charges_b[:pb])
  flows_c = list(flows_a[:pa]) + list(flows_b[:pb])
  mc, dc = diagonal_blocks(charges_c, flows_c, pa)
  out = np.zeros(num_nonzero(charges_c, flows_c), dtype=np.result_type(data_a, data_b))
  for q, pos_a in ma.items():
    # the contracted multi-index set of a's block q carries charge -q under b's (opposite) flows
    if -q in mb and q in mc:
      m, k = da[q]
      nb, kb = db[-q]
      assert k == kb and dc[q] == (m, nb)
      out[mc[q]] = (data_a[pos_a].reshape(m, k) @ data_b[mb[-q]].reshape(nb, k).T).ravel()
  return out, charges_c, flows_c
