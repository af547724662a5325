#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the UNMODIFIED reference (/root/reference/src/{main,errorfunc,log}.cpp) from where it
# lies into oracle/_ref/ (git-ignored; travels to the GPU box like our own built .so files).
# The reference's case switches are hard #defines inside src/main.cpp (:50 TWO_DIMENSIONAL,
# :54 Bar_Module, :55 DAM_Module, :100 MAX_NEIGHBOR_COUNT), so every variant is produced by
# streaming the file through `sed` straight into the compiler's stdin -- no copy of any reference
# source is ever written into this repository.
#
# Products, per variant V in {2d_bar, 2d_dam, 3d_dam, 3d_bar, 2d_turek, 2d_rolling1, 2d_hydro, 2d_rolling2, 2d_rollwall}
# (+ _nb128 big-N flavours):
#   oracle/_ref/Mph_Elastic_Explicit_V      the reference executable (CLI of src/main.cpp:501-508)
#   oracle/_ref/libref_V.so                 same TU + oracle/ref_harness_tail.cpp (stage-level API)
#   oracle/_ref/GeneratorForMph             the reference pre-processor (generator/generator.cpp)
#
# Flags: -O3 -fopenmp (src/makefile:4,7) plus -ffp-contract=off and NO -march=native so that no FMA
# contraction can change bits between builds.
set -euo pipefail
REF=${REF_ROOT:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/src/main.cpp" ]; then
    echo "build_ref.sh: $REF/src/main.cpp not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$OUT"
CXX=${REF_CXX:-g++}   # NOT $CXX: the image exports CXX=/opt/gcc/bin/g++ which has no libgomp
CXXFLAGS="-O3 -fopenmp -ffp-contract=off -Wno-write-strings -Wno-unused-result -w"

variant_sed() {  # $1 = variant name -> sed program on stdout
    local v="$1" prog=""
    local nobar='s|^#define Bar_Module|//#define Bar_Module|'
    case "$v" in
        2d_bar*) prog="" ;;
        2d_dam*) prog="$nobar"';s|^//#define DAM_Module|#define DAM_Module|' ;;
        3d_dam*) prog='s|^#define TWO_DIMENSIONAL|//#define TWO_DIMENSIONAL|;'"$nobar"';s|^//#define DAM_Module|#define DAM_Module|' ;;
        3d_bar*) prog='s|^#define TWO_DIMENSIONAL|//#define TWO_DIMENSIONAL|' ;;
        2d_turek*) prog="$nobar"';s|^//#define Turek_Hron|#define Turek_Hron|' ;;
        2d_rolling1*) prog="$nobar"';s|^//#define Rolling1|#define Rolling1|' ;;
        2d_hydro*) prog="$nobar"';s|^// #define Hydroelastic|#define Hydroelastic|' ;;
        2d_rolling2*) prog="$nobar"';s|^#define DIM 3|#define Rolling2\n#define DIM 3|' ;;   # (no switch line exists for Rolling2)
        2d_rollwall*) prog='s|^#define DIM 3|#define Rolling\n#define DIM 3|' ;;             # Bar_Module + the rolling wall
        *) echo "unknown variant $v" >&2; exit 1 ;;
    esac
    case "$v" in
        *_nb128) prog="$prog;s|^#define MAX_NEIGHBOR_COUNT 512|#define MAX_NEIGHBOR_COUNT 128|" ;;
    esac
    echo "$prog"
}

build_variant() {
    local v="$1"
    local prog; prog="$(variant_sed "$v")"
    local exe="$OUT/Mph_Elastic_Explicit_$v" lib="$OUT/libref_$v.so"
    if [ ! -x "$exe" ] || [ "$REF/src/main.cpp" -nt "$exe" ]; then
        sed -e "$prog" "$REF/src/main.cpp" |
            $CXX $CXXFLAGS -I"$REF/src" -x c++ - "$REF/src/errorfunc.cpp" "$REF/src/log.cpp" -lm -o "$exe"
        echo "built $exe"
    fi
    if [ ! -f "$lib" ] || [ "$HERE/ref_harness_tail.cpp" -nt "$lib" ] || [ "$REF/src/main.cpp" -nt "$lib" ]; then
        { echo '#define main reference_main'; sed -e "$prog" "$REF/src/main.cpp"; echo '#undef main';
          cat "$HERE/ref_harness_tail.cpp"; } |
            $CXX $CXXFLAGS -fPIC -shared -I"$REF/src" -x c++ - "$REF/src/errorfunc.cpp" "$REF/src/log.cpp" -lm -o "$lib"
        echo "built $lib"
    fi
}

VARIANTS=${VARIANTS:-"2d_bar 2d_dam 3d_dam 3d_dam_nb128 2d_turek 2d_rolling1 2d_hydro 2d_rolling2 2d_rollwall"}
for v in $VARIANTS; do build_variant "$v"; done

# the reference pre-processor (generator/makefile builds the same three files)
if [ ! -x "$OUT/GeneratorForMph" ]; then
    $CXX -O2 -w -I"$REF/generator" "$REF/generator/generator.cpp" "$REF/generator/errorfunc.cpp" \
        "$REF/generator/log.cpp" -lm -o "$OUT/GeneratorForMph" 2>/dev/null ||
    $CXX -O2 -w -I"$REF/generator" "$REF/generator/generator.cpp" "$REF/generator/errorfunc.cpp" \
        "$REF/generator/log.cpp" "$REF/generator/typedefs.cpp" -lm -o "$OUT/GeneratorForMph"
    echo "built $OUT/GeneratorForMph"
fi
