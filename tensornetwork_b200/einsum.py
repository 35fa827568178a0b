"""einsum lowered onto the libtnb200 kernels (NumPyBackend.einsum, numpy_backend.py:102-106).

Strategy: (1) per operand, repeated subscripts become a diagonal *view*; subscripts that occur
nowhere else and not in the output are summed away (tnb200_sum); (2) operands are folded left
to right with one (batched) tensordot each: subscripts shared by both operands that are still
needed later or in the output are batch modes, the other shared ones are contracted;
(3) a final transpose view puts the axes in output order."""
from .tensor import B200Tensor


def _parse(expression, nops):
  expression = expression.replace(" ", "")
  if "..." in expression:
    raise NotImplementedError("einsum: ellipsis is not supported by the cuda_b200 backend")
  if "->" in expression:
    lhs, out = expression.split("->")
    ins = lhs.split(",")
  else:
    ins = expression.split(",")
    flat = "".join(ins)
    out = "".join(sorted(c for c in set(flat) if flat.count(c) == 1))
  if len(ins) != nops:
    raise ValueError("einsum: number of operands does not match the subscripts")
  return [list(s) for s in ins], list(out)


def _diag_views(be, t, subs):
  """collapse repeated subscripts of one operand into a diagonal view."""
  while True:
    rep = next((c for c in subs if subs.count(c) > 1), None)
    if rep is None:
      return t, subs
    i = subs.index(rep)
    j = subs.index(rep, i + 1)
    if t.shape[i] != t.shape[j]:
      raise ValueError("einsum: repeated subscript '{}' has unequal extents".format(rep))
    tv = be.torch.diagonal(t.t, dim1=i, dim2=j)  # diagonal axis goes last
    subs = [c for k, c in enumerate(subs) if k not in (i, j)] + [rep]
    t = B200Tensor(tv, t.code)


def einsum(be, expression, *tensors):
  ins, out = _parse(expression, len(tensors))
  ops = []
  for t, subs in zip(tensors, ins):
    t = be.convert_to_tensor(t)
    if len(subs) != t.ndim:
      raise ValueError("einsum: operand has {} axes but subscripts '{}'".format(
          t.ndim, "".join(subs)))
    ops.append(_diag_views(be, t, list(subs)))
  for c in out:
    if not any(c in s for _, s in ops):
      raise ValueError("einsum: output subscript '{}' does not appear in the inputs".format(c))

  def needed_later(c, k):
    return c in out or any(c in s for _, s in ops[k:])

  # sum away subscripts private to one operand and absent from the output
  for k, (t, subs) in enumerate(ops):
    others = [s for j, (_, s) in enumerate(ops) if j != k]
    dead = [i for i, c in enumerate(subs) if c not in out and not any(c in s for s in others)]
    if dead:
      t = be.sum(t, tuple(dead))
      subs = [c for i, c in enumerate(subs) if i not in dead]
      ops[k] = (t, subs)

  acc, asub = ops[0]
  for k in range(1, len(ops)):
    t, subs = ops[k]
    shared = [c for c in asub if c in subs]
    batch = [c for c in shared if needed_later(c, k + 1)]
    contr = [c for c in shared if c not in batch]
    acc = be._contract(acc, t, [asub.index(c) for c in contr], [subs.index(c) for c in contr],  # pylint: disable=protected-access
                       [asub.index(c) for c in batch], [subs.index(c) for c in batch])
    asub = batch + [c for c in asub if c not in shared] + [c for c in subs if c not in shared]
    dead = [i for i, c in enumerate(asub) if not needed_later(c, k + 1)]
    if dead:
      acc = be.sum(acc, tuple(dead))
      asub = [c for i, c in enumerate(asub) if i not in dead]
  if sorted(asub) != sorted(out):
    raise ValueError("einsum: could not reduce to the requested output")
  if asub != out:
    acc = be.transpose(acc, tuple(asub.index(c) for c in out))
  return acc
