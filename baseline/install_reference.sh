#!/bin/bash
# Installs the UNMODIFIED reference package (events555/sdim, pure Python) into baseline/_ref (git-ignored, travels to
# the GPU box with gpurun) so that bench.py can time the reference itself on the box's host cores
# (`cpu_baseline_reference`) and tests/test_integration_stub.py can drive a real reference `Program`.
#
# The prescribed command
#   python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
# fails here: the reference's build backend is poetry-core, which is neither installed nor in /opt/wheelhouse
# (ModuleNotFoundError: No module named 'poetry').  The package is pure Python, so the same pip install is done from a
# scratch copy under /tmp whose [build-system] table names setuptools instead; no file of the package is touched.
# --no-deps: cirq and Diophantine are not installable offline and are not used on the prime-dimension path
# (oracle/ref_harness.py stubs them at import).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${SDIM_REFERENCE_ROOT:-/root/reference}"
[ -d "$REF/sdim" ] || { echo "no reference at $REF"; exit 0; }
TMP="$(mktemp -d /tmp/sdim_ref_XXXX)"
cp -r "$REF/sdim" "$TMP/sdim"
cat > "$TMP/pyproject.toml" <<'P'
[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"
[project]
name = "sdim"
version = "1.2.0"
[tool.setuptools.packages.find]
include = ["sdim*"]
P
rm -rf "$ROOT/baseline/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
       --target "$ROOT/baseline/_ref" "$TMP" 2>&1 | tail -3
rm -rf "$TMP"
# the two shipped circuits the reference's read_circuit resolves relative to its package (sdim/circuit_io.py:6-30)
mkdir -p "$ROOT/baseline/_ref/circuits" && cp "$REF"/circuits/*.chp "$ROOT/baseline/_ref/circuits/" 2>/dev/null || true
ls "$ROOT/baseline/_ref"
