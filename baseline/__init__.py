"""The reference arm: the UNMODIFIED google/TensorNetwork package installed under
baseline/_ref (git-ignored, travels to the GPU box with the gpurun snapshot) plus the import
environment it needs (baseline/refenv.py).  Recipe: tools/install_ref.sh."""
