"""CPU: the N>1 path (subtree partition + point-to-point exchange, replica sharding + gather) with
world_size 2 over gloo."""
import os
import socket
import subprocess
import sys
import pytest
from tensornetwork_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def test_world2_gloo():
  port = str(_free_port())
  procs = []
  for r in range(2):
    env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "parallel_worker.py")], env=env,
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
  outs = [p.communicate(timeout=300)[0] for p in procs]
  assert all(p.returncode == 0 for p in procs), "\n".join(outs)
  assert "PARALLEL OK" in outs[0]


def test_partition_covers_every_step_and_respects_dependencies():
  # caterpillar chain: no parallelism -> critical == total; balanced tree: critical < total
  n = 8
  path = [(0, 1)] + [(0, k) for k in range(n - 2, 0, -1)]   # ((((0,1),2),3)...) chain in linear format
  flops = [1.0] * (n - 1)
  owner, transfers, info = parallel.partition_tree(n, path, flops, 4)
  assert len(owner) == n - 1 and all(o is not None for o in owner)
  assert abs(info["critical"] - info["total"]) < 1e-12
  # balanced: pairs first, then pairs of pairs
  path = [(0, 1), (0, 1), (0, 1), (0, 1), (0, 1), (0, 1), (0, 1)]
  ssa = parallel.path_to_ssa(8, path)
  assert ssa[0] == (0, 1, 8) and ssa[-1][2] == 14
  owner, transfers, info = parallel.partition_tree(8, path, [1.0] * 7, 4)
  assert info["critical"] == 3.0 and info["total"] == 7.0
  assert len(set(owner)) == 4
  for t, src, dst, before in transfers:
    assert src != dst


def _run_workers(script, world, extra_env=None):
  port = str(_free_port())
  procs = []
  for r in range(world):
    env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=port, **(extra_env or {}))
    procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", script)], env=env,
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
  outs = [p.communicate(timeout=300)[0] for p in procs]
  assert all(p.returncode == 0 for p in procs), "\n".join(outs)
  return outs


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_network_executor_over_gloo(world):
  """parallel.ShardedNetwork.run (bench.py's strong_scaling executor) on the world sizes the driver's scaling run uses: result
  equal to the oracle on the root rank, and the point-to-point operations it issues, in order, equal to the model
  (`p2p_issue_order`) the deadlock check simulates.  See tests/sharded_worker.py."""
  outs = _run_workers("sharded_worker.py", world)
  assert "SHARDED OK world=%d" % world in outs[0]


def test_sharded_network_refuses_a_cyclic_plan_on_every_rank():
  outs = _run_workers("sharded_worker.py", 3, {"SHARDED_EXPECT_REFUSAL": "1"})
  assert "SHARDED REFUSED world=3" in outs[0]
