#!/bin/bash
# ORACLE SCAFFOLDING -- builds the UNMODIFIED reference engine from the sources where they lie
# (/root/reference/src) into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
# No reference source is copied into the repo.  Needs: g++, libmpfr.so.6, libgmp.so.10 (runtime
# libraries present in the image; their headers are replaced by oracle/shim/*.h).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${JKQ_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF not present (GPU box) -- using prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
LIBDIR=/usr/lib/x86_64-linux-gnu
CXXFLAGS="-O2 -std=c++14 -w -D_REENTRANT -I$HERE/shim -I$REF"
ENGINE="Simulator QMDDpackage QMDDcomplexD QMDDcircuit QMDDreorder qcost timing textFileUtilities"
FRONT="QASMsimulator QASMscanner QASMtoken"
for f in $ENGINE $FRONT; do
  if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/$f.cpp" -nt "$OUT/obj/$f.o" ]; then
    g++ $CXXFLAGS -c "$REF/$f.cpp" -o "$OUT/obj/$f.o" &
  fi
done
wait
EOBJ=""; for f in $ENGINE; do EOBJ="$EOBJ $OUT/obj/$f.o"; done
FOBJ=""; for f in $FRONT; do FOBJ="$FOBJ $OUT/obj/$f.o"; done
g++ $CXXFLAGS "$HERE/ref_harness.cpp" $EOBJ $LIBDIR/libmpfr.so.6 $LIBDIR/libgmp.so.10 -o "$OUT/jku_ref"
g++ $CXXFLAGS "$HERE/ref_main.cpp" $EOBJ $FOBJ $LIBDIR/libmpfr.so.6 $LIBDIR/libgmp.so.10 -o "$OUT/jku_simulator"
echo "built $OUT/jku_ref $OUT/jku_simulator"
