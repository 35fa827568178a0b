#!/bin/bash
# Install the UNMODIFIED reference (google/TensorNetwork at /root/reference) into baseline/_ref.
# baseline/_ref is git-ignored (no reference source enters history) but NOT gpurun-ignored, so the
# installed package travels to the GPU box, where the -m gpu tests run the reference's own callers
# (tn.Node, tn.ncon, contractors.greedy, split_node*, FiniteDMRG) on backend="cuda_b200" and
# `bench.py --impl reference` times the reference's own numpy backend.
# /root/reference is read-only and setup.py writes build/ + egg-info, so install from a /tmp copy.
# Dependency resolution fails offline (numpy is installed but not in the wheelhouse) => --no-deps.
set -e
cd "$(dirname "$0")/.."
[ -d /root/reference/tensornetwork ] || { echo "no /root/reference here (GPU box?): keeping baseline/_ref as shipped"; exit 0; }
rm -rf /tmp/tn_refcopy baseline/_ref
cp -r /root/reference /tmp/tn_refcopy
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target baseline/_ref /tmp/tn_refcopy 2>&1 | tail -2
rm -rf /tmp/tn_refcopy
python - <<'PY'
from baseline import refenv
tn = refenv.load()
print("reference", tn.__version__, "importable from", refenv.location())
PY
