#!/usr/bin/env bash
# oracle/build_ref.sh — build oracle/_ref/miro_ref from the reference sources WHERE THEY LIE.
#
# Test infrastructure only.  Nothing is copied into the repository: a throw-away
# build directory (oracle/_ref/.build, removed at the end) holds symlinks to the
# reference's files plus four sed-patched translation units generated on the fly.
# We do not run the reference's own build system (src/Makefile is the stale Miro
# base-code Makefile and links GLUT).
#
# Dropped files: Sphere.cpp Triangle.cpp BLPatch.cpp (dead, do not compile),
# MiroWindow.cpp main.cpp (GLUT UI / hard-coded scene needing missing assets),
# Perlin.cpp StoneTexture.cpp (procedural textures, not on the path).
# Patches (all portability, none changes results):
#   ProxyObject.h:14 / ProxyObject.cpp:5  non-const ref bound to a temporary -> const ref
#   hdrloader.cpp:68                      sscanf "%ld" into int (LP64 stack smash) -> "%d"
#   Scene.cpp:297                         count Scene::trace calls (one query = one ray)
#   Texture.cpp:121                       HDR texel reads base+3 past the end -> alpha 1
set -euo pipefail
REF=${MIRO_REFERENCE_SRC:-/root/reference/src}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
B=$OUT/.build
[ -d "$REF" ] || { echo "reference sources not found at $REF" >&2; exit 3; }
rm -rf "$B"; mkdir -p "$B" "$OUT"
DROP="Sphere.cpp Triangle.cpp BLPatch.cpp MiroWindow.cpp main.cpp Perlin.cpp StoneTexture.cpp"
PATCHED="ProxyObject.h ProxyObject.cpp hdrloader.cpp Scene.cpp Texture.cpp"
for f in "$REF"/*.h "$REF"/*.cpp; do
  b=$(basename "$f")
  case " $DROP $PATCHED " in *" $b "*) continue;; esac
  ln -s "$f" "$B/$b"
done
sed 's/Matrix4x4 &t = Matrix4x4()/const Matrix4x4 \&t = Matrix4x4()/' "$REF/ProxyObject.h" > "$B/ProxyObject.h"
sed 's/BVH\* b, Matrix4x4& t)/BVH* b, const Matrix4x4\& t)/' "$REF/ProxyObject.cpp" > "$B/ProxyObject.cpp"
sed 's/-Y %ld +X %ld/-Y %d +X %d/' "$REF/hdrloader.cpp" > "$B/hdrloader.cpp"
sed 's/return m_bvh.intersect(threadID, hitInfo, ray, tMin);/g_miro_trace_calls[threadID*16]++; return m_bvh.intersect(threadID, hitInfo, ray, tMin);/' "$REF/Scene.cpp" > "$B/Scene.cpp"
# only the HDR branch (Texture.cpp:119-121) is rewritten: restore the RGBA branch's 4th channel
python3 - "$REF/Texture.cpp" "$B/Texture.cpp" <<'EOF'
import sys
src = open(sys.argv[1], newline='').read().split('\n')
out, in_hdr = [], False
for line in src:
    if 'm_imageType == HDR' in line: in_hdr = True
    elif 'else' in line and in_hdr and 'HDR' not in line: in_hdr = False
    if in_hdr and 'm_rawData[base + 3]' in line:
        line = line.replace('m_image->m_rawData[base + 3]', '1.0f')
    out.append(line)
open(sys.argv[2], 'w', newline='').write('\n'.join(out))
EOF
grep -q 'const Matrix4x4 &t' "$B/ProxyObject.h" && grep -q 'const Matrix4x4& t' "$B/ProxyObject.cpp" \
  && grep -q -- '-Y %d +X %d' "$B/hdrloader.cpp" && grep -q 'g_miro_trace_calls' "$B/Scene.cpp" \
  || { echo "a reference patch did not apply" >&2; exit 4; }
CXX=${MIRO_CXX:-/usr/bin/g++}   # the image exports CXX=/opt/gcc/bin/g++, which has no libgomp.spec
FLAGS="-std=gnu++14 -fpermissive -w -fopenmp -msse4.1 -O3 -include $HERE/shim/shim.h -I$HERE/shim -I$B -I$HERE/../include"
objs=""
pids=""
for f in "$B"/*.cpp; do
  o="$B/$(basename "$f" .cpp).o"; objs="$objs $o"
  $CXX $FLAGS -c "$f" -o "$o" &
  pids="$pids $!"
done
$CXX $FLAGS -c "$HERE/ref_harness.cpp" -o "$B/ref_harness.o" &
pids="$pids $!"
for p in $pids; do wait "$p"; done
$CXX -fopenmp -O3 $objs "$B/ref_harness.o" -o "$OUT/miro_ref" -lm -ldl
rm -rf "$B"
echo "built $OUT/miro_ref"
