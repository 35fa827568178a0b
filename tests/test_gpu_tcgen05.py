"""GPU parity of the tcgen05/TMA tensor-core path (bf16 / f16 / tf32) against float64 numpy on
identically rounded inputs.  Tolerances: bf16 4e-3 (output rounding 2^-9), f16 1e-3, tf32 1e-3
(10-bit mantissa inputs, fp32 accumulate), all relative Frobenius."""
import numpy as np
import pytest
from util import assert_close, get_backend, rel_err

pytestmark = pytest.mark.gpu
TOLS = {"bfloat16": 4e-3, "float16": 1e-3, "float32": 1e-3}


def _mk(be, rng, shape, dtype):
  x = rng.standard_normal(shape).astype(np.float32)
  if dtype == "float32":
    t = be.convert_to_tensor(x)
  else:
    t = be.astype(be.convert_to_tensor(x), dtype)
  return t, t.to_host().astype(np.float64)


@pytest.mark.parametrize("dtype", ["bfloat16", "float16", "float32"])
@pytest.mark.parametrize("axes", [([2], [0]), ([0], [2]), ([2], [2]), ([0], [0]), ([0, 1], [0, 1]), ([1, 2], [1, 2])])
def test_two_site_all_majors(dtype, axes):
  """A,B (256,2,256): every combination of K-major / MN-major operands (fused transposes)."""
  be = get_backend()
  rng = np.random.default_rng(21)
  A, a = _mk(be, rng, (256, 2, 256), dtype)
  B, b = _mk(be, rng, (256, 2, 256), dtype)
  out = be.tensordot(A, B, axes)
  kern = be.lib.tnb200_last_kernel().decode()
  assert kern.startswith("tcgen05"), kern
  ref = np.tensordot(a, b, axes)
  e = rel_err(out.to_host(), ref)
  assert e < TOLS[dtype], "%s %s via %s: %.3e" % (dtype, axes, kern, e)


@pytest.mark.parametrize("dtype", ["bfloat16", "float32"])
@pytest.mark.parametrize("mkn", [(128, 64, 64), (130, 72, 136), (1000, 520, 264), (64, 1024, 8), (8, 64, 520),
                                 (512, 512, 1024), (2048, 40, 2048), (136, 8, 72)])
def test_ragged_sizes(dtype, mkn):
  be = get_backend()
  rng = np.random.default_rng(22)
  m, k, n = mkn
  A, a = _mk(be, rng, (m, k), dtype)
  B, b = _mk(be, rng, (k, n), dtype)
  out = be.tensordot(A, B, 1)
  kern = be.lib.tnb200_last_kernel().decode()
  assert rel_err(out.to_host(), a @ b) < TOLS[dtype], (kern, mkn)
  # transposed operands: (k,m)^T x (n,k)^T
  At, at = _mk(be, rng, (k, m), dtype)
  Bt, bt = _mk(be, rng, (n, k), dtype)
  out2 = be.tensordot(be.transpose(At), be.transpose(Bt), 1)
  assert rel_err(out2.to_host(), at.T @ bt.T) < TOLS[dtype], (be.lib.tnb200_last_kernel().decode(), mkn)


@pytest.mark.parametrize("dtype", ["bfloat16", "float32"])
def test_batched(dtype):
  be = get_backend()
  rng = np.random.default_rng(23)
  A, a = _mk(be, rng, (6, 256, 128), dtype)
  B, b = _mk(be, rng, (6, 128, 192), dtype)
  out = be.matmul(A, B)
  assert be.lib.tnb200_last_kernel().decode().startswith("tcgen05")
  assert rel_err(out.to_host(), np.matmul(a, b)) < TOLS[dtype]


def test_unaligned_operand_falls_back_to_repack():
  """odd leading dimension -> not TMA addressable -> repacked, still tensor-core, still right."""
  be = get_backend()
  rng = np.random.default_rng(24)
  A, a = _mk(be, rng, (256, 131), "bfloat16")
  B, b = _mk(be, rng, (131, 256), "bfloat16")
  out = be.tensordot(A, B, 1)
  assert rel_err(out.to_host(), a @ b) < TOLS["bfloat16"]


def test_linearity_at_flagship_size():
  """size-independent property at the flagship shape: T(a1 + a2, b) == T(a1, b) + T(a2, b)."""
  be = get_backend()
  rng = np.random.default_rng(25)
  A1, a1 = _mk(be, rng, (512, 2, 512), "float32")
  A2, a2 = _mk(be, rng, (512, 2, 512), "float32")
  B, b = _mk(be, rng, (512, 2, 512), "float32")
  lhs = be.tensordot(be.addition(A1, A2), B, ([2], [0]))
  rhs = be.addition(be.tensordot(A1, B, ([2], [0])), be.tensordot(A2, B, ([2], [0])))
  assert rel_err(lhs.to_host(), rhs.to_host()) < 2e-3


def _launches(be):
  return be.lib.tnb200_launch_count()


@pytest.mark.parametrize("dtype", ["bfloat16", "float32"])
def test_multimode_operands_are_fused_not_repacked(dtype):
  """Operands whose free / contracted group is two non-mergeable modes are addressed in place by
  rank-5 TMA maps: exactly ONE kernel launch (no strided-copy repack), results within tolerance."""
  be = get_backend()
  rng = np.random.default_rng(26)
  # (b) of the cfg-2 zipper: A(512,2,512) x T(2,512,512) over A axes (0,1) <-> T axes (2,0)
  A, a = _mk(be, rng, (512, 2, 512), dtype)
  Tt, t = _mk(be, rng, (2, 512, 512), dtype)
  l0 = _launches(be)
  out = be.tensordot(A, Tt, ([0, 1], [2, 0]))
  assert _launches(be) - l0 == 1, "repacked: %d launches" % (_launches(be) - l0)
  assert be.lib.tnb200_last_kernel().decode().startswith("tcgen05")
  assert rel_err(out.to_host(), np.tensordot(a, t, ([0, 1], [2, 0]))) < TOLS[dtype]
  # free group = two modes around the contracted physical leg (MN-major, inner extent % 64 == 0)
  X, x = _mk(be, rng, (256, 4, 128), dtype)
  Y, y = _mk(be, rng, (4, 192), dtype)
  l0 = _launches(be)
  out = be.tensordot(X, Y, ([1], [0]))
  assert _launches(be) - l0 == 1
  assert rel_err(out.to_host(), np.tensordot(x, y, ([1], [0]))) < TOLS[dtype]
  # K-major operand with two free modes (a slice breaks mergeability), power-of-two inner extent
  Z, z = _mk(be, rng, (8, 64, 256), dtype)
  Zs = Z[:, :32, :]
  W, w = _mk(be, rng, (256, 64), dtype)
  l0 = _launches(be)
  out = be.tensordot(Zs, W, ([2], [0]))
  assert _launches(be) - l0 == 1
  assert rel_err(out.to_host(), np.tensordot(z[:, :32, :], w, ([2], [0]))) < TOLS[dtype]
  # batched + multi-mode
  Ab, ab = _mk(be, rng, (3, 256, 2, 128), dtype)
  Bb, bb = _mk(be, rng, (3, 2, 64, 256), dtype)
  out = be._contract(Ab, Bb, [1, 2], [3, 1], [0], [0])
  ref = np.einsum("bimk,bmji->bkj", ab, bb)
  assert rel_err(out.to_host(), ref) < TOLS[dtype]


@pytest.mark.parametrize("dtype", ["bfloat16", "float32"])
def test_swap_ab_tiny_m(dtype):
  """tiny M under a huge N (the ramp-up steps of the cfg-2 path): computed as C^T tiles, stored transposed."""
  be = get_backend()
  rng = np.random.default_rng(27)
  for (m, k, n) in [(4, 4, 4096), (8, 8, 2048), (16, 16, 1024), (33, 40, 640), (64, 64, 512), (1, 128, 256), (2, 512, 128)]:
    A, a = _mk(be, rng, (m, k), dtype)
    B, b = _mk(be, rng, (k, n), dtype)
    out = be.tensordot(A, B, 1)
    kern = be.lib.tnb200_last_kernel().decode()     # short K + tiny M streams through the CUDA-core kernel
    assert kern.startswith("tcgen05") or kern == "skinny_outer", (m, k, n, kern)
    assert rel_err(out.to_host(), a @ b) < TOLS[dtype], (m, k, n)
    Bt, bt = _mk(be, rng, (n, k), dtype)          # K-major big operand
    out = be.tensordot(A, be.transpose(Bt), 1)
    assert rel_err(out.to_host(), a @ bt.T) < TOLS[dtype], (m, k, n)
  # batched ramp-up step with many small legs: (nb, 4, 2, 2) . (nb, 2, 4, 2, 2, 2, 64, 2, 2) over A[1] <-> B[2]... cfg-2 style
  Ab, ab = _mk(be, rng, (3, 8, 2, 4), dtype)
  Bb, bb = _mk(be, rng, (3, 2, 8, 2, 2, 2, 2, 64, 2, 2), dtype)
  out = be._contract(Ab, Bb, [1], [2], [0], [0])
  ref = np.einsum("bkpq,bxkcdefghi->bpqxcdefghi", ab, bb)
  assert rel_err(out.to_host(), ref) < TOLS[dtype]
