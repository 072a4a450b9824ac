#!/bin/bash
# One-time offline "install" of the reference into baseline/_ref (git-ignored, travels with gpurun).
# The reference is a directory of scripts without setup.py/pyproject, so pip cannot install it:
#   python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
# fails with "does not appear to be a Python project"; we therefore copy the tree verbatim.
set -e
cd "$(dirname "$0")/.."
rm -rf baseline/_ref
if python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /root/reference >/tmp/ref_pip.log 2>&1; then
  echo "pip install succeeded"
else
  echo "pip install failed (expected: no setup.py / pyproject.toml): $(tail -1 /tmp/ref_pip.log)"
  mkdir -p baseline/_ref
  cp -r /root/reference/src baseline/_ref/src
  cp /root/reference/README.md /root/reference/LICENSE baseline/_ref/ 2>/dev/null || true
fi
ls baseline/_ref/src | head -20
