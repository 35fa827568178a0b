"""CPU: the block-sparse index-map builder (pure integer work) reproduces the reference's own maps
bit-exactly (golden vectors from `_find_transposed_diagonal_sparse_blocks`, blocksparse_utils.py:430-634)."""
import numpy as np
from conftest import load_golden
from tensornetwork_b200 import blocksparse as bs


def _legs(z, ci, m):
  return [bs.Index(z["c%d_q%d" % (ci, li)], f) for li, f in enumerate(m["flows"])]


def test_sector_maps_bit_exact():
  meta, z = load_golden("blocksparse")
  for ci, m in enumerate(meta):
    legs = _legs(z, ci, m)
    order = list(range(m["nlegs"])) if m["perm"] is None else list(m["perm"])
    axes_a = m["axes"][0]
    free = [i for i in range(m["nlegs"]) if i not in axes_a]
    order_a = [order[i] for i in free] + [order[i] for i in axes_a]
    qn, dims, maps = bs._sector_maps(legs, order_a, len(free))
    ref_off = z["c%d_mapoff" % ci]
    ref_cat = z["c%d_mapcat" % ci]
    ref_shapes = z["c%d_shapes" % ci]
    ref_q = z["c%d_qnums" % ci]
    assert len(maps) == len(ref_off) - 1, "case %d: number of sectors" % ci
    np.testing.assert_array_equal(dims.T, ref_shapes)
    # reference labels sectors by the fused charge of the row legs (with its own sign convention);
    # ours by the signed row charge: same ordering up to an overall sign
    assert np.array_equal(np.abs(qn), np.abs(ref_q)) or np.array_equal(qn, ref_q) or np.array_equal(qn, -ref_q[::-1])
    for s in range(len(maps)):
      np.testing.assert_array_equal(maps[s], ref_cat[ref_off[s]:ref_off[s + 1]], err_msg="case %d sector %d" % (ci, s))
    assert bs.BlockSparseTensor._nnz(legs) == z["c%d_A" % ci].shape[0]
