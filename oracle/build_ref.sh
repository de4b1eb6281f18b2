#!/bin/bash
# TEST INFRASTRUCTURE (oracle) -- not part of the shipped product.
#
# Builds oracle/_ref/mrt_ref: the reference renderer itself (sources read from
# $MRT_REFERENCE_DIR, default /root/reference; patched in a scratch directory by
# patch_reference.py; nothing of the reference is copied into the repository)
# plus ref_harness.cpp / ref_platform_headless.cpp from this directory.
# Also stages the reference's data assets (earthmap.jpg, obj/*.obj) into the
# git-ignored assets/ directory so that they travel to the GPU box, and decodes
# earthmap.jpg once with the reference's vendored stb_image into assets/earthmap.ppm
# (JPEG decoding is out of scope for the new host).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
SRC="${MRT_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
ASSETS="$ROOT/assets"
if [ ! -d "$SRC" ]; then
    echo "build_ref.sh: $SRC not present; keeping prebuilt $OUT/mrt_ref" >&2
    [ -x "$OUT/mrt_ref" ] && exit 0 || exit 1
fi
mkdir -p "$OUT" "$ASSETS/obj" "$ASSETS/run"
TMP="$(mktemp -d /tmp/mrt_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
python3 "$HERE/patch_reference.py" "$SRC" "$TMP"
cp "$HERE/ref_harness.cpp" "$HERE/ref_platform_headless.cpp" "$TMP/"
# libm canonicalisation (see cr_libm.cpp): separate TU, no builtin folding
g++ -std=c++20 -O2 -ffp-contract=off -fno-builtin -c "$HERE/cr_libm.cpp" -o "$TMP/cr_libm.o"
FILES=""
for f in "$TMP"/*.cpp; do
    case "$(basename "$f")" in
        main.cpp) ;;                      # included by ref_harness.cpp
        *) FILES="$FILES $f" ;;
    esac
done
# -march=x86-64-v3: AVX2 class (mat4.h uses AVX-256); portable to the GPU box's host CPU.
# -ffp-contract=off: no FMA contraction, so results do not depend on the optimiser.
g++ -std=c++20 -O3 -march=x86-64-v3 -ffp-contract=off -fno-exceptions -fpermissive \
    -D__cdecl= -D__stdcall= -w -I"$SRC/include" -I"$TMP" $FILES "$TMP/cr_libm.o" -o "$OUT/mrt_ref" -lpthread -ldl
# CPU-baseline binaries: the reference as it would actually run -- its own flags (clang/clang_build_linux.sh:23-29:
# -std=c++20 -O3 -march=native -fno-exceptions -fno-rtti), the host libm, no interposer, the compiler's default
# floating-point contraction.  Used ONLY for timing (bench.py --impl reference and cpu_baseline); parity uses mrt_ref above.
# -march=native is this container's CPU; the x86-64-v3 twin is the fallback where the GPU box's host CPU lacks an
# instruction set (bench.py probes the native binary first).
TFILES=""
for f in "$TMP"/*.cpp; do
    case "$(basename "$f")" in
        main.cpp) ;;
        *) TFILES="$TFILES $f" ;;
    esac
done
for variant in native x86-64-v3; do
    out="$OUT/mrt_ref_native"; [ "$variant" = "x86-64-v3" ] && out="$OUT/mrt_ref_v3"
    g++ -std=c++20 -O3 -march=$variant -fno-exceptions -fno-rtti -fpermissive -DMRT_REF_TIMING_ONLY \
        -D__cdecl= -D__stdcall= -w -I"$SRC/include" -I"$TMP" $TFILES -o "$out" -lpthread
done
# assets (data, not source)
cp -f "$SRC/earthmap.jpg" "$ASSETS/earthmap.jpg"
for o in bunny.obj Teapot3_no_vt.obj teapot.obj simple.obj pyramid.obj cylinder.obj; do
    cp -f "$SRC/obj/$o" "$ASSETS/obj/$o"
done
# scene.cpp:509 asks for "teapot3_no_vt.obj"; the file is "Teapot3_no_vt.obj"
cp -f "$SRC/obj/Teapot3_no_vt.obj" "$ASSETS/obj/teapot3_no_vt.obj"
chmod -R u+w "$ASSETS"
(cd "$ASSETS/run" && "$OUT/mrt_ref" dump-image ../earthmap.ppm)
echo "built $OUT/mrt_ref"
