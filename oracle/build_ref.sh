#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE, not product code.
#
# Compiles the UNMODIFIED reference sources where they lie under /root/reference
# (never copied into this repo) into oracle/_ref/:
#   libstrelka_ref.so   reference objects + oracle/ref_harness.cpp (a C-ABI shim
#                       that drives the reference's own functions on flattened batches)
#   starling2, strelka2 (optional, `--bins`) the reference binaries, for the demo check
#
# No cmake, no python2, no network: g++ directly on the reference's .cpp files with its
# release flags (-O3 -fomit-frame-pointer -std=c++11, NDEBUG *not* defined; see
# /root/reference/src/cmake/cxxConfigure.cmake:438,452-454), the vendored boost-1.58
# subset (headers + the few compiled libs) and vendored htslib-1.7 (IO structs only).
#
# Outputs only under oracle/_ref/ (git-ignored, but it DOES travel to the GPU box).
set -euo pipefail

HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${STRELKA_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
B="$OUT/build"
JOBS="${JOBS:-$(nproc)}"
WANT_BINS=0
[[ "${1:-}" == "--bins" ]] && WANT_BINS=1

if [[ ! -d "$REF/src/c++/lib" ]]; then
    echo "build_ref.sh: reference tree not found at $REF (expected on the GPU box; using prebuilt oracle/_ref if present)" >&2
    exit 0
fi

mkdir -p "$B/obj" "$B/gen/common"
cd "$B"

# 1. vendored third-party trees ------------------------------------------------------------
for t in boost_1_58_0_subset htslib-1.7-6-g6d2bfb7 rapidjson-1.1.0 CodeMin-1.0.5; do
    [[ -d "$t" ]] || tar xjf "$REF/redist/$t.tar.bz2"
done
if [[ ! -f htslib-1.7-6-g6d2bfb7/libhts.a ]]; then
    (cd htslib-1.7-6-g6d2bfb7 && ./configure --disable-bz2 --disable-lzma --disable-libcurl CFLAGS="-O2 -fPIC" >/dev/null \
        && make -j"$JOBS" lib-static >/dev/null)
fi

# 2. the header cmake would have generated from common/config.h.in -------------------------
cat > gen/common/config.h <<'EOF'
#pragma once
#define WORKFLOW_VERSION "oracle"
#define BUILD_TIME "na"
#define CXX_COMPILER_NAME "g++"
#define COMPILER_VERSION "13"
EOF

INC="-I$B/gen -I$REF/src/c++/lib -I$B/boost_1_58_0_subset -I$B/htslib-1.7-6-g6d2bfb7 -I$B/rapidjson-1.1.0/include -I$B/CodeMin-1.0.5/include"
CXXFLAGS="-std=c++11 -O3 -fomit-frame-pointer -fPIC -w"

# 3. compile list: every non-test library source + the two applications we need -----------
LIBS="alignment appstats assembly blt_common blt_util calibration common errorAnalysis htsapi options starling_common strelka_common"
: > compile.list
for l in $LIBS; do
    find "$REF/src/c++/lib/$l" -name '*.cpp' -not -path '*/test/*' >> compile.list
done
find "$REF/src/c++/lib/applications/strelka" "$REF/src/c++/lib/applications/starling" -name '*.cpp' -not -path '*/test/*' >> compile.list
# test-only helpers the reference's own unit tests link (mock options / IndelBuffer)
find "$REF/src/c++/lib/test" -name '*.cpp' >> compile.list
# boost compiled libs straight from source
for bl in program_options filesystem system timer chrono serialization; do
    find "boost_1_58_0_subset/libs/$bl/src" -name '*.cpp' | grep -v -e windows -e shared_ptr_helper >> compile.list || true
done
find boost_1_58_0_subset/libs/date_time/src/gregorian -name '*.cpp' >> compile.list

objname() { echo "obj/$(echo "$1" | sed -e "s#^$REF/src/c++/##" -e 's#[/+]#_#g' -e 's#\.cpp$#.o#')"; }

cat > Makefile.gen <<EOF
CXX=g++
CXXFLAGS=$CXXFLAGS
INC=$INC
all: objs
EOF
OBJS=""
while read -r f; do
    o="$(objname "$f")"
    OBJS="$OBJS $o"
    extra=""
    case "$f" in *"/lib/test/"*) extra="-include limits";; esac
    printf '%s: %s\n\t@$(CXX) $(CXXFLAGS) $(INC) -I%s %s -c %s -o %s\n' "$o" "$f" "$(dirname "$f")" "$extra" "$f" "$o" >> Makefile.gen
done < compile.list
echo "objs:$OBJS" >> Makefile.gen
make -f Makefile.gen -j"$JOBS" objs

# 4. archives --------------------------------------------------------------------------------
rm -f libcommon.a libapp_strelka.a libapp_starling.a libboost.a libreftest.a
ar rcs libcommon.a $(ls obj/lib_{alignment,appstats,assembly,blt_common,blt_util,calibration,common,errorAnalysis,htsapi,options,starling_common,strelka_common}_*.o)
ar rcs libapp_strelka.a obj/lib_applications_strelka_*.o
ar rcs libapp_starling.a obj/lib_applications_starling_*.o
ar rcs libreftest.a obj/lib_test_*.o
ar rcs libboost.a obj/boost_1_58_0_subset_*.o

# 5. the harness .so --------------------------------------------------------------------------
g++ $CXXFLAGS $INC -include limits -I"$REF/src/c++/lib/starling_common" -c "$HERE/ref_harness.cpp" -o obj/ref_harness.o
g++ $CXXFLAGS $INC -include limits -I"$REF/src/c++/lib/starling_common" -c "$HERE/ref_harness_score_indels.cpp" -o obj/ref_harness_score_indels.o
g++ -shared -o "$OUT/libstrelka_ref.so" obj/ref_harness.o obj/ref_harness_score_indels.o \
    -Wl,--start-group libapp_strelka.a libreftest.a libcommon.a -Wl,--end-group libboost.a \
    htslib-1.7-6-g6d2bfb7/libhts.a -lz -lpthread
echo "built $OUT/libstrelka_ref.so"

# 6. optional: the reference binaries (demo / end-to-end checks) -----------------------------
if [[ $WANT_BINS == 1 ]]; then
    for app in starling2 strelka2; do
        g++ $CXXFLAGS $INC -c "$REF/src/c++/bin/$app.cpp" -o obj/main_$app.o
    done
    g++ -o "$OUT/starling2" obj/main_starling2.o -Wl,--start-group libapp_starling.a libcommon.a -Wl,--end-group libboost.a htslib-1.7-6-g6d2bfb7/libhts.a -lz -lpthread
    g++ -o "$OUT/strelka2"  obj/main_strelka2.o  -Wl,--start-group libapp_strelka.a  libcommon.a -Wl,--end-group libboost.a htslib-1.7-6-g6d2bfb7/libhts.a -lz -lpthread
    echo "built $OUT/starling2 $OUT/strelka2"
fi
