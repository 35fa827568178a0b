"""numpy restatement of the reference's contraction drivers (test oracle).

* `ncon`            — tensornetwork/ncon_interface.py:523-663 + `_jittable_ncon` :364-520
* `greedy_path`     — the path provider used on both sides of every parity test
                      (numpy's `_greedy_path`; reproduces the three greedy known-answers of
                      contractors/opt_einsum_paths/path_calculation_test.py:83-93)
* `contract_path`   — contractors/opt_einsum_paths/path_contractors.py:36-97 (`base`) with
                      `contract_between` network_components.py:1984-2095 semantics
All arithmetic goes through oracle.np_backend.
"""
import numpy as np
from . import np_backend as nb


# ----------------------------------------------------------------------------- ncon
def canonicalize(network_structure):
  """ncon_interface.py:69-115 — map labels to +/- integers, order-preserving."""
  flat = [l for sub in network_structure for l in sub]
  neg_int = sorted({l for l in flat if not isinstance(l, str) and l < 0})
  pos_int = sorted({l for l in flat if not isinstance(l, str) and l > 0})
  neg_str = sorted({l for l in flat if isinstance(l, str) and l[0] == '-'},
                   reverse=True)
  pos_str = sorted({l for l in flat if isinstance(l, str) and l[0] != '-'})
  mapping = dict(zip(neg_str + neg_int,
                     range(-len(neg_int + neg_str), 0)))
  mapping.update(dict(zip(pos_int + pos_str,
                          range(1, 1 + len(pos_int + pos_str)))))
  return [[mapping[l] for l in labels] for labels in network_structure], mapping


def _partial_trace(tensor, labels):
  """ncon_interface.py:241-277 — transpose -> reshape -> trace."""
  trace_labels = [l for l in labels if labels.count(l) == 2]
  if not trace_labels:
    return tensor, labels, []
  num_cont = len(trace_labels) // 2
  uniq = sorted(trace_labels)[0:-1:2]
  pos = [[n for n, l in enumerate(labels) if l == t] for t in uniq]
  contracted = [p[0] for p in pos] + [p[1] for p in pos]
  free = [n for n in range(len(labels)) if n not in contracted]
  shape = tensor.shape
  cdim = int(np.prod([shape[d] for d in contracted[:num_cont]]))
  tmp = tuple([shape[p] for p in free] + [cdim, cdim])
  res = nb.trace(nb.reshape(nb.transpose(tensor, tuple(free + contracted)), tmp))
  return res, [l for l in labels if l not in uniq], uniq


def _batch_cont(t1, t2, tensors, net, con_order, common_batch, l1, l2):
  """ncon_interface.py:280-354 — transpose -> reshape -> matmul -> reshape."""
  common_batch = list(common_batch)
  b1 = [l1.index(l) for l in common_batch]
  b2 = [l2.index(l) for l in common_batch]
  nb1 = {l for l in l1 if l not in common_batch}
  nb2 = {l for l in l2 if l not in common_batch}
  cc = list(nb1.intersection(nb2))
  c1 = [l1.index(l) for l in cc]
  c2 = [l2.index(l) for l in cc]
  f1 = set(l1) - set(cc) - set(common_batch)
  f2 = set(l2) - set(cc) - set(common_batch)
  fp1 = [n for n, l in enumerate(l1) if l in f1]
  fp2 = [n for n, l in enumerate(l2) if l in f2]
  s1 = np.array(t1.shape)
  s2 = np.array(t2.shape)
  ns1 = (np.prod(s1[b1]), np.prod(s1[fp1]), np.prod(s1[c1]))
  ns2 = (np.prod(s2[b2]), np.prod(s2[c2]), np.prod(s2[fp2]))
  m1 = nb.reshape(nb.transpose(t1, tuple(b1 + fp1 + c1)), ns1)
  m2 = nb.reshape(nb.transpose(t2, tuple(b2 + c2 + fp2)), ns2)
  res = nb.matmul(m1, m2)
  final = tuple(np.concatenate([s1[b1], s1[fp1], s2[fp2]]).astype(int))
  res = nb.reshape(res, final)
  net.append([l1[i] for i in b1] + [l1[i] for i in fp1] + [l2[i] for i in fp2])
  tensors.append(res)
  con_order = [c for c in con_order if c not in cc]
  return tensors, net, con_order


def ncon(tensors, network_structure, con_order=None, out_order=None):
  """ncon_interface.py:523-663 / 364-520 on numpy arrays."""
  tensors = [np.asarray(t) for t in tensors]
  if out_order == []:
    out_order = None
  if con_order == []:
    con_order = None
  net, mapping = canonicalize(network_structure)
  flat = [l for sub in net for l in sub]
  uniq = list(set(flat))
  if out_order is None:
    out_order = sorted([l for l in uniq if l < 0], reverse=True)
  else:
    out_order = [mapping[o] for o in out_order]
  if con_order is None:
    con_order = sorted([l for l in uniq if l > 0])
  else:
    con_order = [mapping[o] for o in con_order]
  init_con_order = list(con_order)

  for n, t in enumerate(tensors):
    tensors[n], net[n], contracted = _partial_trace(t, net[n])
    if contracted:
      con_order = [c for c in con_order if c not in contracted]
  flat = [l for sub in net for l in sub]
  single = [l for l in flat if flat.count(l) == 1 and l > 0]
  if single:
    con_order = [o for o in con_order if o not in single]
  for loc, labels in enumerate(net):
    if set(labels).intersection(single):
      inds = [labels.index(l) for l in single if l in labels]
      net[loc] = [l for l in labels if l not in single]
      tensors[loc] = nb.tsum(tensors[loc], tuple(inds))

  skip = 0
  batch_labels, batch_cnts = [], []
  for l in set(flat):
    cnt = flat.count(l)
    if cnt > 2 or (cnt == 2 and l < 0):
      batch_labels.append(l)
      batch_cnts.append(cnt)

  while con_order:
    ci = con_order[0]
    if ci in batch_labels:
      con_order.append(con_order.pop(0))
      skip += 1
      if skip > len(con_order):
        raise ValueError("ncon seems stuck in an infinite loop; con_order = "
                         "{}".format(init_con_order))
      continue
    locs = [n for n, labels in enumerate(net) if ci in labels]
    t2 = tensors.pop(locs[1])
    t1 = tensors.pop(locs[0])
    l2 = net.pop(locs[1])
    l1 = net.pop(locs[0])
    common = list(set(l1).intersection(l2))
    c1 = [l1.index(l) for l in common]
    c2 = [l2.index(l) for l in common]
    cb = set(batch_labels).intersection(common)
    if cb:
      delete = []
      for i, bl in enumerate(batch_labels):
        if bl in cb:
          batch_cnts[i] -= 1
          if bl > 0 and batch_cnts[i] <= 2:
            delete.append(i)
          elif bl < 0 and batch_cnts[i] < 2:
            delete.append(i)
      for i in sorted(delete, reverse=True):
        del batch_cnts[i]
        del batch_labels[i]
      tensors, net, con_order = _batch_cont(t1, t2, tensors, net, con_order, cb,
                                            l1, l2)
    else:
      srt = [c1.index(l) for l in sorted(c1)]
      tensors.append(nb.tensordot(t1, t2, (tuple(c1[i] for i in srt),
                                           tuple(c2[i] for i in srt))))
      net.append([l for l in l1 if l not in common] +
                 [l for l in l2 if l not in common])
      con_order = [c for c in con_order if c not in common]

  while len(tensors) > 1:
    t2 = tensors.pop()
    t1 = tensors.pop()
    l2 = net.pop()
    l1 = net.pop()
    common = list(set(l1).intersection(l2))
    cb = set(batch_labels).intersection(common)
    if cb:
      tensors, net, con_order = _batch_cont(t1, t2, tensors, net, con_order, cb,
                                            l1, l2)
    else:
      tensors.append(nb.outer_product(t1, t2))
      net.append(l1 + l2)

  if len(net[0]) > 1:
    labels = net[0]
    return nb.transpose(tensors[0], tuple(labels.index(l) for l in out_order))
  return tensors[0]


# ------------------------------------------------------------- path contraction
def greedy_path(labels, out_labels, size_dict, memory_limit=None):
  """Greedy pairwise order [(i, j), ...] in opt_einsum's `path` convention.

  Stand-in for `opt_einsum.paths.greedy` (third-party, requirements.txt:3, not
  vendored, not installed): numpy's own greedy einsum path search.
  """
  from numpy._core.einsumfunc import _greedy_path  # pylint: disable=import-outside-toplevel
  input_sets = [set(l) for l in labels]
  return [tuple(p) for p in _greedy_path(
      input_sets, set(out_labels), dict(size_dict),
      2**62 if memory_limit is None else memory_limit)]


def contract_between(t1, l1, t2, l2):
  """network_components.py:2048-2085 without an output_edge_order:
  axes of t1 sorted ascending, output = free(t1) + free(t2)."""
  shared = [l for l in l1 if l in l2]
  if not shared:
    return nb.outer_product(t1, t2), list(l1) + list(l2)
  a1 = [l1.index(l) for l in shared]
  a2 = [l2.index(l) for l in shared]
  srt = [a1.index(x) for x in sorted(a1)]
  a1 = [a1[i] for i in srt]
  a2 = [a2[i] for i in srt]
  out = nb.tensordot(t1, t2, [a1, a2])
  return out, [l for l in l1 if l not in shared] + [l for l in l2 if l not in shared]


def contract_path(tensors, labels, path, out_labels):
  """path_contractors.py:86-96 — pop (a, b), append the pair's product, final
  reorder to `out_labels` (reorder_edges -> transpose, network_components.py:246).
  Labels appearing twice on one tensor (trace edges) are not handled here; the
  callers in this repo never build them for the path contractors."""
  tensors = [np.asarray(t) for t in tensors]
  labels = [list(l) for l in labels]
  for a, b in path:
    t, l = contract_between(tensors[a], labels[a], tensors[b], labels[b])
    for i in sorted([a, b], reverse=True):
      del tensors[i]
      del labels[i]
    tensors.append(t)
    labels.append(l)
  res, lab = tensors[0], labels[0]
  if len(lab) > 1:
    res = nb.transpose(res, tuple(lab.index(l) for l in out_labels))
  return res


def network_flops(labels, path, size_dict):
  """Algorithmic work of a path: list of (M, K, N) per pairwise step (SURVEY 8d:
  flops = 2MNK, bytes = (MK+KN+MN)*sizeof)."""
  labels = [list(l) for l in labels]
  steps = []
  for a, b in path:
    l1, l2 = labels[a], labels[b]
    shared = [l for l in l1 if l in l2]
    K = int(np.prod([size_dict[l] for l in shared])) if shared else 1
    M = int(np.prod([size_dict[l] for l in l1 if l not in shared] or [1]))
    N = int(np.prod([size_dict[l] for l in l2 if l not in shared] or [1]))
    steps.append((M, K, N))
    new = [l for l in l1 if l not in shared] + [l for l in l2 if l not in shared]
    for i in sorted([a, b], reverse=True):
      del labels[i]
    labels.append(new)
  return steps
