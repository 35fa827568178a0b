import json
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
  config.addinivalue_line("markers", "refhost: needs /root/reference (build container only)")


def load_golden(name):
  z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
  meta = json.loads(str(z["__meta__"]))
  return meta, z


@pytest.fixture(scope="session")
def golden():
  return load_golden
