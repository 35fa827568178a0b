"""Runs in a subprocess (build container only): the REAL reference's callers on backend="cuda_b200",
with the device layer replaced by tests/fake_lib.FakeLib (host memory).  Checks the adapter's host
logic and the registration path against the reference's own numpy backend."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_shim
tn = ref_shim.load()                       # reference first, so the adapter subclasses the real AbstractBackend
from tensornetwork_b200 import _lib, backend as tb_backend
import fake_lib
_lib.set_lib(fake_lib.FakeLib())
tb_backend._CONFIG["device"] = "cpu"
import tensornetwork_b200 as tb
assert tb.registered and tb_backend.HAVE_TENSORNETWORK
from tensornetwork.backends import backend_factory, abstract_backend
be = backend_factory.get_backend("cuda_b200")
assert isinstance(be, abstract_backend.AbstractBackend) and be.name == "cuda_b200"
assert backend_factory.get_backend("cuda_b200") is be   # singleton per name (backend_factory.py:42-46)


def host(x):
  return np.asarray(x)


rng = np.random.default_rng(0)
# ---- Node @ Node, contract_between, reorder (network_components.py:1984-2095)
a, b = rng.standard_normal((4, 5, 6)), rng.standard_normal((6, 5, 3))
na, nb_ = tn.Node(a, backend="cuda_b200"), tn.Node(b, backend="cuda_b200")
na[2] ^ nb_[0]
na[1] ^ nb_[1]
c = na @ nb_
np.testing.assert_allclose(host(c.tensor), np.tensordot(a, b, ([2, 1], [0, 1])), atol=1e-12)
# trace edge + flatten
t = rng.standard_normal((3, 4, 3, 5))
nt = tn.Node(t, backend="cuda_b200")
nt[0] ^ nt[2]
np.testing.assert_allclose(host(tn.contract_trace_edges(nt).tensor), np.trace(t, axis1=0, axis2=2), atol=1e-12)
# outer product, conj, scalar multiplication through Node operators
x, y = rng.standard_normal((2, 3)), rng.standard_normal((4,))
np.testing.assert_allclose(host(tn.outer_product(tn.Node(x, backend="cuda_b200"), tn.Node(y, backend="cuda_b200")).tensor),
                           np.tensordot(x, y, 0), atol=1e-12)

# ---- ncon on the reference's own implementation, several structures (ncon_interface.py:523)
for ts, net, con, out in [
    ([rng.standard_normal((10, 10)), rng.standard_normal((10, 10))], [(-1, 1), (1, -2)], None, None),
    ([rng.standard_normal((3, 4, 4)), rng.standard_normal((3, 5))], [(1, 2, 2), (1, -1)], None, None),
    ([rng.standard_normal((7, 3, 4)), rng.standard_normal((7, 4, 5))], [(-1, -2, 1), (-1, 1, -3)], None, None),
    ([rng.standard_normal((2, 3)), rng.standard_normal((4, 5))], [(-1, -2), (-3, -4)], None, [-3, -1, -4, -2]),
    ([rng.standard_normal((3, 4, 5)), rng.standard_normal((5, 4, 6)), rng.standard_normal((6, 3))],
     [(1, 2, 3), (3, 2, 4), (4, 1)], [3, 2, 4, 1], None),
]:
  ref = tn.ncon(ts, net, con_order=con, out_order=out, backend="numpy")
  got = tn.ncon(ts, net, con_order=con, out_order=out, backend="cuda_b200")
  np.testing.assert_allclose(host(got), ref, atol=1e-12)

# ---- contractors.greedy on a small <psi|psi> (path_contractors.py:165)
L, D = 6, 4
dims = [1] + [min(D, 2**min(i, L - i)) for i in range(1, L)] + [1]
kets = [rng.standard_normal((dims[i], 2, dims[i + 1])) for i in range(L)]


def build(backend):
  k = [tn.Node(x, backend=backend) for x in kets]
  bnodes = [tn.Node(np.conj(x), backend=backend) for x in kets]
  for i in range(L):
    k[i][1] ^ bnodes[i][1]
    if i + 1 < L:
      k[i][2] ^ k[i + 1][0]
      bnodes[i][2] ^ bnodes[i + 1][0]
  k[0][0] ^ bnodes[0][0]
  k[-1][2] ^ bnodes[-1][2]
  return k + bnodes


r_np = tn.contractors.greedy(build("numpy")).tensor
r_cu = tn.contractors.greedy(build("cuda_b200")).tensor
np.testing.assert_allclose(host(r_cu), r_np, rtol=1e-12)

# ---- split_node family (network_operations.py:130-588)
t4 = rng.standard_normal((4, 5, 6, 3))
for kw in ({}, {"max_singular_values": 7}, {"max_truncation_err": 0.5, "relative": True}):
  n1, n2 = tn.Node(t4, backend="numpy"), tn.Node(t4, backend="cuda_b200")
  l1, r1, e1 = tn.split_node(n1, [n1[0], n1[1]], [n1[2], n1[3]], **kw)
  l2, r2, e2 = tn.split_node(n2, [n2[0], n2[1]], [n2[2], n2[3]], **kw)
  assert host(l2.tensor).shape == l1.tensor.shape and host(e2).shape == np.asarray(e1).shape
  np.testing.assert_allclose(np.tensordot(host(l2.tensor), host(r2.tensor), 1), np.tensordot(l1.tensor, r1.tensor, 1), atol=1e-10)
n2 = tn.Node(t4, backend="cuda_b200")
u, s, vh, _ = tn.split_node_full_svd(n2, [n2[0], n2[1]], [n2[2], n2[3]], max_singular_values=5)
assert host(s.tensor).shape == (5, 5)
n2 = tn.Node(t4, backend="cuda_b200")
q, r = tn.split_node_qr(n2, [n2[0], n2[1]], [n2[2], n2[3]])
np.testing.assert_allclose(np.tensordot(host(q.tensor), host(r.tensor), 1), t4, atol=1e-12)
n2 = tn.Node(t4, backend="cuda_b200")
r, q = tn.split_node_rq(n2, [n2[0], n2[1]], [n2[2], n2[3]])
np.testing.assert_allclose(np.tensordot(host(r.tensor), host(q.tensor), 1), t4, atol=1e-12)

# ---- error conventions (numpy_backend.py:92-97, :41) and default-backend machinery
try:
  be.convert_to_tensor([1, 2])
  raise SystemExit("expected TypeError")
except TypeError:
  pass
try:
  be.tensordot(be.convert_to_tensor(np.ones((2, 3))), be.convert_to_tensor(np.ones((4, 5))), [[1], [0]])
  raise SystemExit("expected ValueError")
except ValueError:
  pass
tn.set_default_backend("cuda_b200")
assert tn.Node(np.ones(3)).backend.name == "cuda_b200"
tn.set_default_backend("numpy")

# ---- two-site DMRG on the reference's FiniteDMRG (dmrg.py:445), XXZ N=6: energy vs the numpy backend
if "--dmrg" in sys.argv:
  N, Dm = 6, 8

  def energy(backend):
    np.random.seed(10)
    mps = tn.FiniteMPS.random([2] * N, [Dm] * (N - 1), dtype=np.float64, backend=backend)
    mpo = tn.FiniteXXZ(np.ones(N - 1), np.ones(N - 1), np.zeros(N), dtype=np.float64, backend=backend)
    dmrg = tn.FiniteDMRG(mps, mpo)
    return float(np.real(np.asarray(dmrg.run_two_site(max_bond_dim=Dm, num_sweeps=4, num_krylov_vecs=10, verbose=0))))
  e_np, e_cu = energy("numpy"), energy("cuda_b200")
  assert abs(e_np - e_cu) < 1e-8, (e_np, e_cu)
  print("dmrg energies", e_np, e_cu)
print("REFHOST OK")
