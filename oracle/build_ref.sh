#!/bin/bash
#  TEST INFRASTRUCTURE (oracle/) -- not part of the shipped product.
#
#  Compiles the UNMODIFIED reference `meryl` (marbl/meryl + meryl-utility) from
#  the sources where they lie under $REF (default /root/reference) into
#  oracle/_ref/meryl.  Nothing is copied into this repository: objects, the
#  generated version.H and the binary all go under oracle/_ref/ (git-ignored).
#
#  The reference's own build system is not used (it needs htslib's external
#  libraries); instead the count/print/statistics/union path is compiled from
#  its own source files with g++, with the BAM/CRAM reader replaced by
#  oracle/hts_stub.C.
#
#  usage: oracle/build_ref.sh [-f]     (no-op when $REF is absent: the GPU
#                                       box uses the prebuilt oracle/_ref/meryl)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
SRC="$REF/src"

if [ ! -d "$SRC/meryl" ]; then
  echo "build_ref: $SRC not present; keeping prebuilt $OUT/meryl" >&2
  exit 0
fi
if [ -x "$OUT/meryl" ] && [ -x "$OUT/meryl-lookup" ] && [ "$1" != "-f" ]; then
  exit 0
fi

mkdir -p "$OUT/obj" "$OUT/gen"
( cd /tmp && perl "$SRC/utility/src/version.pl" meryl release v1.4.2 "$OUT/gen/version.H" > /dev/null )

CXX="${REF_CXX:-/usr/bin/g++}"   #  not $CXX: this image exports a wrapper without libgomp
FLAGS="-std=c++2a -fopenmp -pthread -O3 -DNOBACKTRACE -mxsave -w \
  -I$OUT/gen -I$SRC -I$SRC/meryl -I$SRC/meryl-lookup -I$SRC/utility/src -I$SRC/utility/src/sequence"

FILES=""
for d in bits datastructures files kmers-v1 math system; do
  FILES="$FILES $(ls $SRC/utility/src/$d/*.C)"
done
FILES="$FILES $SRC/utility/src/sequence/dnaSeq-v1.C $SRC/utility/src/sequence/bufSeqFile-v1.C $SRC/utility/src/sequence/sequence-v1.C"
FILES="$FILES $HERE/hts_stub.C"
NUTIL=$(echo $FILES | wc -w)
#  the two programs: meryl (count / print / statistics / set operations) and meryl-lookup (existence / include / exclude)
FILES="$FILES $(ls $SRC/meryl/*.C)"
LOOKUP="$SRC/meryl-lookup/meryl-lookup.C $SRC/meryl-lookup/meryl-lookup-help.C $SRC/meryl-lookup/dump.C $SRC/meryl-lookup/existence.C $SRC/meryl-lookup/include-exclude.C"
FILES="$FILES $LOOKUP"

pids=()
objs=""
lobjs=""
n=0
for f in $FILES; do
  o="$OUT/obj/$(echo "$f" | sed -e 's#/#_#g' -e 's#\.C$#.o#')"
  if [ $n -lt $NUTIL ]; then objs="$objs $o"; lobjs="$lobjs $o";
  elif echo "$f" | grep -q "/meryl-lookup/"; then lobjs="$lobjs $o";
  else objs="$objs $o"; fi
  $CXX $FLAGS -c "$f" -o "$o" &
  n=$((n+1))
  if [ $((n % $(nproc))) -eq 0 ]; then wait; fi
done
wait
$CXX -fopenmp -pthread -O3 -o "$OUT/meryl" $objs -lm
$CXX -fopenmp -pthread -O3 -o "$OUT/meryl-lookup" $lobjs -lm
rm -rf "$OUT/obj"
echo "build_ref: built $OUT/meryl and $OUT/meryl-lookup" >&2
