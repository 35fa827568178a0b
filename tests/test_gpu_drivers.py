"""GPU parity of the drivers (ncon / greedy path / helpers) vs golden vectors + oracle."""
import numpy as np
import pytest
from conftest import load_golden
from util import assert_close, get_backend
from oracle import np_backend as nb
from oracle import np_network as nn

pytestmark = pytest.mark.gpu


def test_golden_ncon():
  from tensornetwork_b200 import drivers
  be = get_backend()
  meta, z = load_golden("ncon")
  for i, m in enumerate(meta):
    ts = [z["c%d_t%d" % (i, j)] for j in range(m["n"])]
    out = drivers.ncon(ts, m["net"], m["con"], m["out"], backend=be)
    assert_close(out, z["c%d_out" % i], what="ncon case %d" % i)


def test_golden_greedy():
  from tensornetwork_b200 import drivers
  be = get_backend()
  meta, z = load_golden("greedy")
  for ci, m in enumerate(meta):
    if m.get("open"):
      ts = [z["open_a"], z["open_b"], z["open_c"]]
      out, ref = m["out"], z["open_out"]
    else:
      kets = [z["c%d_k%d" % (ci, j)] for j in range(m["L"])]
      ts = kets + [np.conj(k) for k in kets]
      out, ref = [], z["c%d_out" % ci]
    res = drivers.contract_network(ts, m["labels"], out, backend=be)
    assert_close(res, ref, what="greedy case %d" % ci)


def test_elementwise_helpers():
  be = get_backend()
  rng = np.random.default_rng(11)
  for dt in ("float64", "float32", "complex128"):
    x = rng.standard_normal((5, 6, 7)).astype(dt)
    y = rng.standard_normal((5, 6, 7)).astype(dt)
    if dt.startswith("complex"):
      x = x + 1j * rng.standard_normal((5, 6, 7))
      y = y - 1j * rng.standard_normal((5, 6, 7))
    X, Y = be.convert_to_tensor(x), be.convert_to_tensor(y)
    assert_close(be.addition(X, Y), x + y)
    assert_close(be.subtraction(X, Y), x - y)
    assert_close(be.multiply(X, Y), x * y)
    assert_close(be.divide(X, Y), x / y)
    assert_close(be.conj(X), np.conj(x))
    assert_close(be.sqrt(be.abs(X)), np.sqrt(np.abs(x)))
    assert_close(be.exp(X), np.exp(x))
    assert_close(be.sign(X), np.sign(x))
    assert_close(be.norm(X), np.asarray(np.linalg.norm(x)))
    assert_close(be.sum(X, (0, 2)), np.sum(x, axis=(0, 2)))
    assert_close(be.sum(X, (1,), keepdims=True), np.sum(x, axis=(1,), keepdims=True))
    assert_close(X * 2.5, x * 2.5)
    assert_close(3.0 - X, 3.0 - x)
    assert_close(be.transpose(X, (2, 0, 1)) + be.transpose(Y, (2, 0, 1)), np.transpose(x + y, (2, 0, 1)))
    v = rng.standard_normal(7).astype(x.real.dtype)
    assert_close(be.broadcast_right_multiplication(X, be.convert_to_tensor(v)), x * v)
    w = rng.standard_normal(5).astype(x.real.dtype)
    assert_close(be.broadcast_left_multiplication(be.convert_to_tensor(w), X), x * w[:, None, None])
    Z = be.copy(X)
    Z /= be.norm(X)
    assert_close(Z, x / np.linalg.norm(x))
  m = rng.standard_normal((4, 6, 6, 3))
  M = be.convert_to_tensor(m)
  assert_close(be.trace(M, axis1=1, axis2=2), np.trace(m, axis1=1, axis2=2))
  assert_close(be.trace(M, offset=2, axis1=1, axis2=2), np.trace(m, offset=2, axis1=1, axis2=2))
  assert_close(be.diagflat(be.convert_to_tensor(m[0, 0])), np.diagflat(m[0, 0]))
  assert_close(be.diagflat(be.convert_to_tensor(m[0, 0, 0]), k=-2), np.diagflat(m[0, 0, 0], k=-2))
  assert_close(be.diagonal(M, axis1=1, axis2=2), np.diagonal(m, axis1=1, axis2=2))
  assert_close(be.eye(4, M=6), np.eye(4, M=6))
  assert_close(be.ones((2, 3), np.float32), np.ones((2, 3), np.float32))
  assert_close(be.outer_product(be.convert_to_tensor(m[0, 0]), be.convert_to_tensor(m[1, 1])), np.tensordot(m[0, 0], m[1, 1], 0))
  r = be.randn((256, 256), np.float64, seed=3).to_host()
  assert abs(r.mean()) < 0.02 and abs(r.std() - 1) < 0.02
  assert be.reshape(be.transpose(M, (3, 1, 0, 2)), (18, 24)).shape == (18, 24)
  assert_close(be.reshape(be.transpose(M, (3, 1, 0, 2)), (18, 24)), np.reshape(np.transpose(m, (3, 1, 0, 2)), (18, 24)))
  with pytest.raises(ValueError):
    be.broadcast_right_multiplication(M, M)


def test_einsum():
  be = get_backend()
  rng = np.random.default_rng(12)
  a, b = rng.standard_normal((3, 4, 5)), rng.standard_normal((4, 3, 6))
  # numpy_backend_test.py:132-138: 'ij,jil->l'
  x, y = rng.standard_normal((2, 3)), rng.standard_normal((3, 2, 4))
  assert_close(be.einsum("ij,jil->l", be.convert_to_tensor(x), be.convert_to_tensor(y)), np.einsum("ij,jil->l", x, y))
  assert_close(be.einsum("ijk,jil->kl", be.convert_to_tensor(a), be.convert_to_tensor(b)), np.einsum("ijk,jil->kl", a, b))
  assert_close(be.einsum("ijk,jil->ikl", be.convert_to_tensor(a), be.convert_to_tensor(b)), np.einsum("ijk,jil->ikl", a, b))
  c = rng.standard_normal((4, 4, 3))
  assert_close(be.einsum("iij->j", be.convert_to_tensor(c)), np.einsum("iij->j", c))
  # CopyNode-style (network_components.py:903-908)
  p, q, r = rng.standard_normal((3, 2)), rng.standard_normal((3, 4)), rng.standard_normal((3, 5))
  assert_close(be.einsum("ia,ib,ic->abc", *[be.convert_to_tensor(t) for t in (p, q, r)]), np.einsum("ia,ib,ic->abc", p, q, r))


def test_lanczos_vs_golden():
  be = get_backend()
  meta, z = load_golden("lanczos")
  for i, m in enumerate(meta):
    h = be.convert_to_tensor(z["h%d" % i])
    x0 = be.convert_to_tensor(z["x%d" % i])
    ev, vecs = be.eigsh_lanczos(lambda x, mat: be.tensordot(mat, x, ([1], [0])), [h], x0,
                                num_krylov_vecs=m["num_krylov_vecs"], numeig=m["numeig"],
                                reorthogonalize=m["reorthogonalize"], ndiag=m["ndiag"])
    np.testing.assert_allclose(ev, z["ev%d" % i], rtol=1e-9, atol=1e-9)
    for j, v in enumerate(vecs):
      ref = z["vec%d" % i][j]
      got = v.to_host()
      s = np.sign(np.vdot(ref, got))
      np.testing.assert_allclose(got * s, ref, rtol=0, atol=1e-7)


def test_cfg2_mps_norm_small_sizes_and_properties():
  """<psi|psi> greedy contraction (cfg 2 shape family) vs the oracle at a size the oracle
  finishes quickly, fp64 <= 1e-10."""
  from tensornetwork_b200 import drivers
  be = get_backend()
  rng = np.random.default_rng(3)
  L, D, d = 16, 64, 2
  dims = [1] + [min(D, d**min(i, L - i)) for i in range(1, L)] + [1]
  kets = [rng.standard_normal((dims[i], d, dims[i + 1])) / np.sqrt(dims[i] * d) for i in range(L)]
  labels = []
  for side in "kb":
    for i in range(L):
      labels.append(["e0" if i == 0 else "%s%d" % (side, i), "p%d" % i, "eL" if i == L - 1 else "%s%d" % (side, i + 1)])
  ts = kets + [np.conj(k) for k in kets]
  sizes = {l: t.shape[ax] for t, labs in zip(ts, labels) for ax, l in enumerate(labs)}
  path = nn.greedy_path(labels, [], sizes)
  ref = nn.contract_path(ts, labels, path, [])
  out = drivers.contract_network(ts, labels, [], path=path, backend=be)
  assert_close(out, ref, tol=1e-10)


def _norm_network(L, D, d=2):
  dims = [1] + [min(D, d**min(i, L - i)) for i in range(1, L)] + [1]
  labels = []
  for side in "kb":
    for i in range(L):
      labels.append(["e0" if i == 0 else "%s%d" % (side, i), "p%d" % i, "eL" if i == L - 1 else "%s%d" % (side, i + 1)])
  return dims, labels


# float32 runs on the tensor cores as TF32 (2^-11 input rounding per pairwise step); the ket and bra halves make the
# same rounding errors, so over the ~2L dependent steps they add coherently: stated tolerance 3e-2 on the scalar
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 3e-2)])
def test_compiled_network_graph_replay_staging_and_aliases(dtype, tol):
  """CUDA-graph CompiledNetwork (what bench.py times): batched samples, pinned staging arena, bra layer as
  conj-alias views of the ket buffers; every replay must equal the oracle's pairwise contraction per sample."""
  import torch
  from tensornetwork_b200 import drivers
  be = get_backend()
  rng = np.random.default_rng(13)
  L, D, NB = 12, 32, 3
  dims, labels = _norm_network(L, D)
  core = [(dims[i], 2, dims[i + 1]) for i in range(L)] * 2
  shapes = [(NB,) + s for s in core]
  sizes = {l: s[ax] for s, labs in zip(core, labels) for ax, l in enumerate(labs)}
  path = nn.greedy_path(labels, [], sizes)
  net = drivers.CompiledNetwork(be, shapes, np.dtype(dtype), labels, [], path=path, nbatch=1,
                                conj_aliases={L + i: i for i in range(L)})
  host = net.host_staging()
  assert all(host[L + i] is None for i in range(L))
  for rep in range(2):                                 # second replay with fresh data through the same graph
    kets = [(rng.standard_normal((NB,) + core[i]) / np.sqrt(core[i][0] * 2)).astype(dtype) for i in range(L)]
    for i in range(L):
      host[i].copy_(torch.from_numpy(kets[i]))
    out = net.run_staged().to_host()
    for b in range(NB):
      ts = [k[b].astype(np.float64) for k in kets]
      ref = nn.contract_path(ts + [np.conj(t) for t in ts], labels, path, [])
      assert abs(out[b] - ref) <= tol * abs(ref), (dtype, rep, b, out[b], ref)
  # unbatched graph fed from device tensors via load(); explicit (non-aliased) bra inputs
  net1 = drivers.CompiledNetwork(be, core, np.dtype(dtype), labels, [], path=path, nbatch=0)
  ts = [k[0] for k in kets]
  dev = [be.convert_to_tensor(t) for t in ts + [np.conj(t) for t in ts]]
  out1 = net1(dev).to_host()
  ref = nn.contract_path([t.astype(np.float64) for t in ts] + [np.conj(t).astype(np.float64) for t in ts], labels, path, [])
  assert abs(float(out1) - ref) <= tol * abs(ref)
  assert net1.launches_per_replay >= len(path)


@pytest.mark.parametrize("dtype,tol", [("bfloat16", 6e-2), ("float32", 3e-2)])
def test_chained_launch_equals_stepwise(dtype, tol, monkeypatch):
  """The MPS zipper of a D=256 norm network as ONE chained persistent launch (tnb200_chain_*) must reproduce the
  step-by-step graph bit for bit (same tiles, same k order, same rounding) and the oracle within the dtype's
  tolerance; replays must be stable (dependency counters are reset by every launch)."""
  import torch
  from tensornetwork_b200 import drivers
  monkeypatch.setenv("TNB200_CHAIN_FORCE", "1")    # 5 samples are below the size at which chaining pays off
  monkeypatch.setenv("TNB200_CHAIN_G", "2")        # several rounds + short producer->consumer distance
  be = get_backend()
  rng = np.random.default_rng(17)
  L, D, NB = 24, 256, 5
  dims, labels = _norm_network(L, D)
  core = [(dims[i], 2, dims[i + 1]) for i in range(L)] * 2
  shapes = [(NB,) + s for s in core]
  sizes = {l: s[ax] for s, labs in zip(core, labels) for ax, l in enumerate(labs)}
  path = nn.greedy_path(labels, [], sizes)
  kets = [(rng.standard_normal((NB,) + core[i]) / np.sqrt(core[i][0] * 2)).astype(np.float32) for i in range(L)]
  dev = [be.astype(be.convert_to_tensor(k), dtype) for k in kets]
  al = {L + i: i for i in range(L)}
  net_c = drivers.CompiledNetwork(be, shapes, dtype, labels, [], path=path, nbatch=1, conj_aliases=al, use_chains=True)
  net_s = drivers.CompiledNetwork(be, shapes, dtype, labels, [], path=path, nbatch=1, conj_aliases=al, use_chains=False)
  assert net_c.chains and max(len(c.steps) for c in net_c.chains) >= 4, "no chain was formed"
  assert net_c.launches_per_replay < net_s.launches_per_replay
  net_c.load(dev + [None] * L)
  net_s.load(dev + [None] * L)
  ref = net_s().to_host().astype(np.float64)
  for rep in range(3):
    out = net_c().to_host().astype(np.float64)
    if dtype == "bfloat16":
      np.testing.assert_array_equal(out, ref)          # same tiles, same k order, same rounding
    else:
      np.testing.assert_allclose(out, ref, rtol=2e-6)   # the step-by-step plan may pick other tile shapes in fp32
  for b in range(NB):
    ts = [d.to_host()[b].astype(np.float64) for d in dev]
    exact = nn.contract_path(ts + [np.conj(t) for t in ts], labels, path, [])
    assert abs(out[b] - exact) <= tol * abs(exact), (dtype, b, out[b], exact)


def test_cfg2_full_size_fp64_vs_oracle_and_scaling_property():
  """BASELINE configs[1] at FULL size (L=64, D=512, d=2, 127 pairwise contractions): one network in fp64 against the
  oracle's pairwise contraction along the same greedy path (<= 1e-10), then the size-independent property
  <c psi|c psi> = c^2 <psi|psi> on the bf16 CUDA-graph path the bench times: scaling one site tensor by 2 (exact in
  bf16) must scale every sample's result by exactly 4 — bit for bit, through the chained launch."""
  from tensornetwork_b200 import drivers
  be = get_backend()
  rng = np.random.default_rng(3)
  L, D = 64, 512
  dims, labels = _norm_network(L, D)
  core = [(dims[i], 2, dims[i + 1]) for i in range(L)] * 2
  sizes = {l: s[ax] for s, labs in zip(core, labels) for ax, l in enumerate(labs)}
  path = nn.greedy_path(labels, [], sizes)
  assert len(path) == 127
  kets = [rng.standard_normal(core[i]) / np.sqrt(core[i][0] * 2) for i in range(L)]
  ref = nn.contract_path(kets + [np.conj(k) for k in kets], labels, path, [])
  out = drivers.contract_network([be.convert_to_tensor(k) for k in kets] + [be.convert_to_tensor(np.conj(k)) for k in kets],
                                 labels, [], path=path, backend=be)
  assert abs(float(out.to_host()) - ref) <= 1e-10 * abs(ref)
  # bf16, batched samples, graph + chain: exact power-of-two scaling
  NB = 4
  shapes = [(NB,) + s for s in core]
  net = drivers.CompiledNetwork(be, shapes, "bfloat16", labels, [], path=path, nbatch=1,
                                conj_aliases={L + i: i for i in range(L)})
  dev = [be.astype(be.convert_to_tensor((rng.standard_normal((NB,) + core[i]) / np.sqrt(core[i][0] * 2)).astype(np.float32)),
                   "bfloat16") for i in range(L)]
  net.load(dev + [None] * L)
  base = net().to_host().astype(np.float64).copy()
  assert np.all(np.isfinite(base)) and np.all(base > 0)
  for site in (0, 31, 63):
    scaled = list(dev)
    scaled[site] = type(dev[site])(dev[site].t * 2, dev[site].code)       # exact in bf16
    net.load(scaled + [None] * L)
    np.testing.assert_array_equal(net().to_host().astype(np.float64), 4.0 * base)
  # and within bf16 tolerance of the fp64 oracle on the same (bf16-rounded) inputs, sample 0
  ts = [d.to_host()[0].astype(np.float64) for d in dev]
  exact = nn.contract_path(ts + [np.conj(t) for t in ts], labels, path, [])
  # 127 steps with bf16-rounded intermediates (2^-9 relative each, ket and bra halves coherent): stated tolerance 3e-2,
  # checked on every sample of the batch (measured ~1e-2)
  for b in range(NB):
    tb_ = [d.to_host()[b].astype(np.float64) for d in dev]
    ex = nn.contract_path(tb_ + [np.conj(t) for t in tb_], labels, path, [])
    assert abs(base[b] - ex) <= 3e-2 * abs(ex), (b, base[b], ex)


def test_sharded_network_single_rank_equals_oracle():
  """parallel.ShardedNetwork with world = 1 (no transfers): the whole tree is one local subtree replayed as a CUDA graph;
  result equals the numpy oracle along the same path (fp64, 1e-10).  The N > 1 schedule is exercised by the gloo tests
  (host logic) and by `bench.py --gpus N` (`strong_scaling.parity_ok`)."""
  import sys, os
  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  import bench
  from tensornetwork_b200 import drivers, parallel
  be = get_backend()
  labels, sizes, shapes, dims = bench.ttn_network({"b3": 48, "b2": 12, "b1": 6, "p": 4})
  path = drivers.greedy_path(labels, [], sizes)
  rng = np.random.default_rng(9)
  n_ket = len(labels) // 2
  kets = [rng.standard_normal(shapes[i]) / np.sqrt(np.prod(shapes[i][1:])) for i in range(n_ket)]
  host = kets + [np.conj(k) for k in kets]
  ref = float(nn.contract_path(host, labels, path, []))
  sh = parallel.ShardedNetwork(be, shapes, np.float64, labels, path, 0, 1)
  sh.load([be.convert_to_tensor(h) for h in host])
  for _ in range(2):
    out, root = sh.run()
  assert root == 0 and abs(float(out.to_host()) - ref) <= 1e-10 * abs(ref)
  # the partition used at N = 4 keeps a ket half and its bra half on one rank: only small tensors cross ranks
  labels_f, sizes_f, shapes_f, _ = bench.ttn_network(None)          # the benchmark's dimensions, symbolically
  path_f = drivers.greedy_path(labels_f, [], sizes_f)
  flops_f = [2.0 * m * k * n_ for m, k, n_ in nn.network_flops(labels_f, path_f, sizes_f)]
  owner, transfers, info = parallel.partition_tree(len(labels_f), path_f, flops_f, 4)
  ssa = parallel.path_to_ssa(len(labels_f), path_f)
  lab = {i: list(l) for i, l in enumerate(labels_f)}
  for a, b, o in ssa:
    shared = [l for l in lab[a] if l in lab[b]]
    lab[o] = [l for l in lab[a] if l not in shared] + [l for l in lab[b] if l not in shared]
  biggest = max(int(np.prod([sizes_f[l] for l in lab[t]])) for t, _, _, _ in transfers)
  assert biggest <= 64 * 64 * 16 and max(info["per_rank"]) <= 1.05 * info["total"] / 4
