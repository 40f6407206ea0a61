#!/usr/bin/env bash
# Test infrastructure only (never used by the product path).
#
# Compiles the reference's three native extension modules *from the sources where
# they lie* under /root/reference (nothing is copied into the repo) into
# oracle/_ref/*.so.  Mirrors /root/reference/discretize/_extensions/meson.build:33-63
# (cython -> C/C++17 -> shared module).  oracle/refload.py then imports the
# reference's pure-python package from /root/reference and resolves
# discretize._extensions.{interputils_cython,tree_ext,simplex_helpers} to these
# binaries.  Only usable where /root/reference exists (the build container).
set -euo pipefail
REF=${REF_ROOT:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
BUILD="$OUT/build"
EXT="$REF/discretize/_extensions"
[ -d "$EXT" ] || { echo "reference not present at $REF; nothing to build"; exit 0; }
mkdir -p "$BUILD"
PY=${PYTHON:-python}
SUFFIX=$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')
PYINC=$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')
NPINC=$($PY -c 'import numpy;print(numpy.get_include())')
CFLAGS="-O2 -fPIC -shared -w -DNPY_NO_DEPRECATED_API=NPY_1_22_API_VERSION -I$PYINC -I$NPINC -I$EXT"

if [ ! -f "$OUT/interputils_cython$SUFFIX" ]; then
  $PY -m cython -3 -Xfreethreading_compatible=True "$EXT/interputils_cython.pyx" -o "$BUILD/interputils_cython.c"
  gcc $CFLAGS "$BUILD/interputils_cython.c" -o "$OUT/interputils_cython$SUFFIX" -lm
fi
if [ ! -f "$OUT/simplex_helpers$SUFFIX" ]; then
  $PY -m cython -3 --cplus -Xfreethreading_compatible=True "$EXT/simplex_helpers.pyx" -o "$BUILD/simplex_helpers.cpp"
  g++ -std=c++17 $CFLAGS "$BUILD/simplex_helpers.cpp" -o "$OUT/simplex_helpers$SUFFIX"
fi
if [ ! -f "$OUT/tree_ext$SUFFIX" ]; then
  $PY -m cython -3 --cplus -Xfreethreading_compatible=True "$EXT/tree_ext.pyx" -o "$BUILD/tree_ext.cpp"
  g++ -std=c++17 $CFLAGS "$BUILD/tree_ext.cpp" "$EXT/tree.cpp" "$EXT/geom.cpp" -o "$OUT/tree_ext$SUFFIX"
fi
rm -rf "$BUILD"
ls -la "$OUT"
