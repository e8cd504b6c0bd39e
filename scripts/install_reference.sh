#!/bin/bash
# Puts the UNMODIFIED reference (interestingLSY/swiftLLM) under baseline/_ref/ so that scripts/ref_triton_bench.py can time its
# own Triton path on a B200 next to bench.py (north_star: "... next to the reference's own Triton path on one B200").
#
# The reference installs itself with `pip install -e .` + `pip install -e csrc` (its README, "Build and Run"): an editable install
# exposes the source tree itself, and its setup.py depends on that (packages=["swiftllm"] leaves out swiftllm.server / swiftllm.worker,
# so a regular `pip install --target` yields a package that cannot be imported).  The offline equivalent of the editable install is
# what this script does: a verbatim copy of the tree at baseline/_ref/src (git-ignored, travels with gpurun) that the bench script
# puts on sys.path, and an in-place build of the `swiftllm_c` extension (host-only C++, builds without a GPU).
# baseline/_ref is NOT part of the product: nothing under swiftllm_b200/, tests/ or bench.py reads it.
set -euo pipefail
cd "$(dirname "$0")/.."
REF=${1:-/root/reference}
[ -d "$REF/swiftllm" ] || { echo "reference tree not found at $REF"; exit 1; }
rm -rf baseline/_ref
mkdir -p baseline/_ref
cp -r "$REF" baseline/_ref/src
cd baseline/_ref/src/csrc
TORCH_CUDA_ARCH_LIST=10.0a MAX_JOBS=4 python setup.py build_ext --inplace > ../../build_swiftllm_c.log 2>&1 || { tail -20 ../../build_swiftllm_c.log; exit 1; }
rm -rf build   # intermediate objects (the .so next to setup.py is what gets imported)
ls -la swiftllm_c*.so
echo "reference ready under baseline/_ref/src"
